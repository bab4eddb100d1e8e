# GPU session F (round 2): 16-worker-warp attention (forward schedule 5, backward with 4 column groups): numerics + timing +
# step A/B; DRAM traffic of every kernel of one step (-> profiles/r02_gemm_traffic.json); cfg3 launch list; full suite.
set -x
O=gpurun_out/r2f
mkdir -p $O
B200_ATTN_FWD_SCHEDULE=5 timeout 200 python tools/attn_check.py fwd > $O/attn_fwd_s5.log 2>&1; echo "rc=$?" >> $O/attn_fwd_s5.log
B200_ATTN_BWD_GROUPS=4 timeout 200 python tools/attn_check.py bwd > $O/attn_bwd_g4.log 2>&1; echo "rc=$?" >> $O/attn_bwd_g4.log
B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 200 python tools/attn_check.py time > $O/attn_time_s5_g4.log 2>&1; echo "rc=$?" >> $O/attn_time_s5_g4.log
B200_ATTN_FWD_SCHEDULE=1 B200_ATTN_BWD_GROUPS=2 timeout 200 python tools/attn_check.py time > $O/attn_time_s1_g2.log 2>&1; echo "rc=$?" >> $O/attn_time_s1_g2.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_ATTN_BWD_GROUPS=4 timeout 300 python bench.py $Q > $O/bench_bwd4.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 timeout 300 python bench.py $Q > $O/bench_fwd5.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 300 python bench.py $Q > $O/bench_fwd5_bwd4.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q > $O/bench_fwd5_bwd4_packed.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu_s5g4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_s5g4.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
# DRAM bytes + duration of every launch of one cfg2 step (2 passes per kernel), then one cfg3 step (durations only)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/traffic_step_cfg2.csv python tools/profile_step.py > $O/traffic_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step_cfg3.csv python tools/profile_step.py 0 cfg3 > $O/launches_cfg3.log 2>&1
# source-level capture of the 16-warp attention kernels
B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_.*tc" -s 8 -c 4 -o $O/attn_tc16 python tools/attn_check.py time > $O/ncu_attn_tc16.log 2>&1
ls -la $O; cat $O/attn_time_s5_g4.log $O/attn_time_s1_g2.log; tail -3 $O/attn_fwd_s5.log; tail -4 $O/attn_bwd_g4.log; tail -n 4 $O/pytest_gpu.log; tail -n 4 $O/pytest_gpu_s5g4.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'])"; done
