# GPU session H (round 2): attention with mask-free chunk bodies (forward schedule 5, backward 2 / 4 column groups);
# teacher forward on a second stream; full suite.
set -x
O=gpurun_out/r2h
mkdir -p $O
B200_ATTN_FWD_SCHEDULE=5 timeout 200 python tools/attn_check.py fwd > $O/attn_fwd_s5.log 2>&1; echo "rc=$?" >> $O/attn_fwd_s5.log
B200_ATTN_BWD_GROUPS=4 timeout 200 python tools/attn_check.py bwd > $O/attn_bwd_g4.log 2>&1; echo "rc=$?" >> $O/attn_bwd_g4.log
B200_ATTN_BWD_GROUPS=2 timeout 200 python tools/attn_check.py bwd > $O/attn_bwd_g2.log 2>&1; echo "rc=$?" >> $O/attn_bwd_g2.log
B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 200 python tools/attn_check.py time > $O/attn_time_s5_g4.log 2>&1; echo "rc=$?" >> $O/attn_time_s5_g4.log
B200_ATTN_FWD_SCHEDULE=1 B200_ATTN_BWD_GROUPS=2 timeout 200 python tools/attn_check.py time > $O/attn_time_s1_g2.log 2>&1; echo "rc=$?" >> $O/attn_time_s1_g2.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_ATTN_BWD_GROUPS=4 timeout 300 python bench.py $Q > $O/bench_bwd4.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 timeout 300 python bench.py $Q > $O/bench_fwd5.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 300 python bench.py $Q > $O/bench_fwd5_bwd4.json 2>> $O/bench_ab.err
B200_TEACHER_STREAM=1 timeout 300 python bench.py $Q > $O/bench_tstream.json 2>> $O/bench_ab.err
B200_TEACHER_STREAM=1 timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -x > $O/pytest_tstream.log 2>&1; echo "pytest rc=$?" >> $O/pytest_tstream.log
cat $O/attn_time_s5_g4.log $O/attn_time_s1_g2.log; tail -3 $O/attn_fwd_s5.log; tail -4 $O/attn_bwd_g4.log; tail -4 $O/attn_bwd_g2.log; tail -n 4 $O/pytest_gpu.log; tail -n 3 $O/pytest_tstream.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'])"; done
