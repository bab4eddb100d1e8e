# GPU session U (round 2): last full validation of the committed tree: suite, smoke, default bench line.
set -x
O=gpurun_out/r2u
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
tail -n 4 $O/pytest_gpu.log; tail -n 2 $O/smoke.log; tail -n 2 $O/bench_cfg2.err
python -c "import json; d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['loss_delta_vs_oracle'], d['gpu_torch_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'])"
