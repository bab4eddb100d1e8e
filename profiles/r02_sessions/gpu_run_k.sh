# GPU session K (round 2): state of the tree with the final defaults -- suite, smoke, full bench lines (cfg2 with every baseline,
# cfg3 / cfg4 / cfg5), launch list of one step, source-level ncu of the attention kernels (forward schedule 8, backward).
set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
timeout 600 python bench.py --config cfg3 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "rc=$?" >> $O/bench_cfg3.err
timeout 600 python bench.py --config cfg5 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
timeout 600 python bench.py --config cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "rc=$?" >> $O/bench_cfg4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step_cfg2.csv python tools/profile_step.py > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tc" -s 4 -c 2 -o $O/attn_bwd python tools/attn_check.py time > $O/ncu_attn_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc8" -s 4 -c 2 -o $O/attn_fwd8 python tools/attn_check.py time > $O/ncu_attn_fwd8.log 2>&1
ls -la $O; tail -n 4 $O/pytest_gpu.log; tail -n 3 $O/smoke.log
for f in $O/bench_cfg*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d.get('parity') and d['parity'].get('loss_delta_vs_oracle'), (d.get('gpu_torch_baseline') or {}).get('value'))"; done
