# GPU session D (round 2): attention forward schedules (1 / 3 / 4) and the shared-memory CE kernel, numerics + timing; ncu
# source-level captures of the tcgen05 attention kernels; full suite.  Outputs -> gpurun_out/r2d/
set -x
O=gpurun_out/r2d
mkdir -p $O
for sc in 4 3 1; do
  B200_ATTN_FWD_SCHEDULE=$sc timeout 200 python tools/attn_check.py fwd > $O/attn_fwd_s$sc.log 2>&1; echo "rc=$?" >> $O/attn_fwd_s$sc.log
  B200_ATTN_FWD_SCHEDULE=$sc timeout 200 python tools/attn_check.py time > $O/attn_time_s$sc.log 2>&1; echo "rc=$?" >> $O/attn_time_s$sc.log
done
timeout 200 python tools/attn_check.py bwd > $O/attn_bwd.log 2>&1; echo "rc=$?" >> $O/attn_bwd.log
for v in 0 4; do B200_CE_VARIANT=$v timeout 200 python tools/loss_bench.py >> $O/loss_bench.log 2>&1; done
B200_CE_VARIANT=4 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q > $O/pytest_ce4.log 2>&1; echo "rc=$?" >> $O/pytest_ce4.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_CE_VARIANT=4 timeout 300 python bench.py $Q > $O/bench_ce4.json 2>> $O/bench_ab.err
B200_CE_VARIANT=4 B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=4 timeout 300 python bench.py $Q > $O/bench_ce4_fwd4.json 2>> $O/bench_ab.err
B200_CE_VARIANT=4 B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=4 B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q > $O/bench_ce4_fwd4_packed.json 2>> $O/bench_ab.err
# source-level captures of the tcgen05 attention kernels (few launches; --import-source needs -lineinfo: built with it)
B200_ATTN_FWD_SCHEDULE=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_.*tc" -s 8 -c 4 -o $O/attn_tc python tools/attn_check.py time > $O/ncu_attn_tc.log 2>&1
B200_CE_VARIANT=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dino_ce" -s 4 -c 3 -o $O/ce4 python tools/loss_bench.py > $O/ncu_ce4.log 2>&1
ls -la $O; for sc in 4 3 1; do grep time $O/attn_time_s$sc.log; tail -2 $O/attn_fwd_s$sc.log; done; grep variant $O/loss_bench.log; tail -n 6 $O/pytest_gpu.log; tail -n 3 $O/pytest_ce4.log
