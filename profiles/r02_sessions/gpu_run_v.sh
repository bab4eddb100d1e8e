# GPU session V (round 2): the tree after removing the measured-and-rejected attention schedules: suite, smoke, attention check,
# default bench line.
set -x
O=gpurun_out/r2v
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 200 python tools/attn_check.py time > $O/attn_time.log 2>&1
B200_ATTN_FWD_SCHEDULE=1 timeout 200 python tools/attn_check.py fwd > $O/attn_fwd_s1.log 2>&1
timeout 900 python bench.py --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
tail -n 4 $O/pytest_gpu.log; tail -n 2 $O/smoke.log; cat $O/attn_time.log; tail -3 $O/attn_fwd_s1.log
python -c "import json; d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['loss_delta_vs_oracle'])"
