# GPU session J (round 2): forward schedule 8 (persistent schedule 7 with the next tile's Q / K / V prefetched).
set -x
O=gpurun_out/r2j
mkdir -p $O
B200_ATTN_FWD_SCHEDULE=8 timeout 200 python tools/attn_check.py fwd > $O/attn_fwd_s8.log 2>&1; echo "rc=$?" >> $O/attn_fwd_s8.log
B200_ATTN_FWD_SCHEDULE=8 timeout 200 python tools/attn_check.py time > $O/attn_time_s8.log 2>&1; echo "rc=$?" >> $O/attn_time_s8.log
B200_ATTN_FWD_SCHEDULE=7 timeout 200 python tools/attn_check.py time > $O/attn_time_s7.log 2>&1; echo "rc=$?" >> $O/attn_time_s7.log
B200_ATTN_FWD_SCHEDULE=8 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x > $O/pytest_s8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_s8.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default_s7.json 2>> $O/bench_ab.err
B200_ATTN_FWD_SCHEDULE=8 timeout 300 python bench.py $Q > $O/bench_s8.json 2>> $O/bench_ab.err
B200_ATTN_FWD_SCHEDULE=8 B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q > $O/bench_s8_packed.json 2>> $O/bench_ab.err
timeout 300 python bench.py $Q --config cfg5 > $O/bench_cfg5_s7.json 2>> $O/bench_ab.err
B200_ATTN_FWD_SCHEDULE=8 timeout 300 python bench.py $Q --config cfg5 > $O/bench_cfg5_s8.json 2>> $O/bench_ab.err
cat $O/attn_time_s8.log $O/attn_time_s7.log; tail -9 $O/attn_fwd_s8.log; tail -n 3 $O/pytest_s8.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['loss'])"; done
