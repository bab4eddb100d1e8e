# GPU session N (round 2): packed short-sequence backward with one tile per CTA and two CTAs per SM (local crops).
set -x
O=gpurun_out/r2n
mkdir -p $O
timeout 200 python tools/attn_check.py bwd > $O/attn_bwd.log 2>&1; echo "rc=$?" >> $O/attn_bwd.log
timeout 200 python tools/attn_check.py time > $O/attn_time.log 2>&1; echo "rc=$?" >> $O/attn_time.log
B200_TC_ATTN_PACKED=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -x > $O/pytest_packed.log 2>&1; echo "pytest rc=$?" >> $O/pytest_packed.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q > $O/bench_packed.json 2>> $O/bench_ab.err
timeout 300 python bench.py $Q --config cfg5 > $O/bench_cfg5_default.json 2>> $O/bench_ab.err
B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q --config cfg5 > $O/bench_cfg5_packed.json 2>> $O/bench_ab.err
cat $O/attn_time.log; grep "N=37\|N=54\|N=16" $O/attn_bwd.log; tail -n 3 $O/pytest_packed.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['loss'])"; done
