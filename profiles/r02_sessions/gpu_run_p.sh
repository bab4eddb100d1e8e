# GPU session P (round 2): final defaults (packed tcgen05 attention for the local crops included) -- suite, smoke, bench lines,
# launch list.
set -x
O=gpurun_out/r2p
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
timeout 600 python bench.py --config cfg3 --no-cpu-baseline --no-gpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "rc=$?" >> $O/bench_cfg3.err
timeout 600 python bench.py --config cfg5 --no-cpu-baseline --no-gpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" >> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step_cfg2.csv python tools/profile_step.py > $O/launches.log 2>&1
timeout 200 python tools/attn_check.py time > $O/attn_time.log 2>&1
ls -la $O; tail -n 4 $O/pytest_gpu.log; tail -n 3 $O/smoke.log; cat $O/attn_time.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), d.get('parity') and d['parity'].get('loss_delta_vs_oracle'), (d.get('gpu_torch_baseline') or {}).get('value'))"; done
