# GPU session L (round 2): backward schedule 2 (key-tile-outer blocks, fused softmax-backward pass): numerics, timing, step A/B.
set -x
O=gpurun_out/r2l
mkdir -p $O
B200_ATTN_BWD_SCHEDULE=2 timeout 200 python tools/attn_check.py bwd > $O/attn_bwd_s2.log 2>&1; echo "rc=$?" >> $O/attn_bwd_s2.log
B200_ATTN_BWD_SCHEDULE=2 timeout 200 python tools/attn_check.py time > $O/attn_time_b2.log 2>&1; echo "rc=$?" >> $O/attn_time_b2.log
B200_ATTN_BWD_SCHEDULE=2 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > $O/pytest_b2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_b2.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_ATTN_BWD_SCHEDULE=2 timeout 300 python bench.py $Q > $O/bench_b2.json 2>> $O/bench_ab.err
timeout 300 python bench.py $Q --config cfg5 > $O/bench_cfg5_default.json 2>> $O/bench_ab.err
B200_ATTN_BWD_SCHEDULE=2 timeout 300 python bench.py $Q --config cfg5 > $O/bench_cfg5_b2.json 2>> $O/bench_ab.err
cat $O/attn_time_b2.log; tail -12 $O/attn_bwd_s2.log; tail -n 3 $O/pytest_b2.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['loss'])"; done
