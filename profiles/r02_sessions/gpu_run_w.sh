# GPU session W (round 2): final tree (after the loss.cu / schedule clean-ups): suite, smoke, loss microbench, bench lines.
set -x
O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 200 python tools/loss_bench.py > $O/loss_bench.log 2>&1
timeout 900 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
timeout 600 python bench.py --config cfg3 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "rc=$?" >> $O/bench_cfg3.err
timeout 600 python bench.py --config cfg5 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
timeout 600 python bench.py --config cfg4 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "rc=$?" >> $O/bench_cfg4.err
tail -n 4 $O/pytest_gpu.log; tail -n 2 $O/smoke.log; cat $O/loss_bench.log
for f in $O/bench_cfg*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), d.get('parity') and d['parity'].get('loss_delta_vs_oracle'), (d.get('gpu_torch_baseline') or {}).get('value'))"; done
