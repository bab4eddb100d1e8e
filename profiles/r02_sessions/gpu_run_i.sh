# GPU session I (round 2): forward schedule 7 (P in tensor memory, A operand of P V from TMEM, two CTAs per SM): numerics,
# timing, step A/B, suite; source-level ncu of it and of the fp32-residual GEMM epilogue.
set -x
O=gpurun_out/r2i
mkdir -p $O
B200_ATTN_FWD_SCHEDULE=7 timeout 200 python tools/attn_check.py fwd > $O/attn_fwd_s7.log 2>&1; echo "rc=$?" >> $O/attn_fwd_s7.log
B200_ATTN_FWD_SCHEDULE=7 B200_ATTN_BWD_GROUPS=4 timeout 200 python tools/attn_check.py time > $O/attn_time_s7_g4.log 2>&1; echo "rc=$?" >> $O/attn_time_s7_g4.log
B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 200 python tools/attn_check.py time > $O/attn_time_s5_g4.log 2>&1; echo "rc=$?" >> $O/attn_time_s5_g4.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=5 B200_ATTN_BWD_GROUPS=4 timeout 300 python bench.py $Q > $O/bench_fwd5_bwd4.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=7 B200_ATTN_BWD_GROUPS=4 timeout 300 python bench.py $Q > $O/bench_fwd7_bwd4.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=7 B200_ATTN_BWD_GROUPS=4 B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q > $O/bench_fwd7_bwd4_packed.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_ATTN_FWD_SCHEDULE=7 B200_ATTN_BWD_GROUPS=4 timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_s7g4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_s7g4.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B200_ATTN_FWD_SCHEDULE=7 B200_ATTN_BWD_GROUPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_.*tc" -s 8 -c 2 -o $O/attn_tc7 python tools/attn_check.py time > $O/ncu_attn_tc7.log 2>&1
CASES=gemm timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tcgen05" -o $O/gemm_cases python tools/gemm_cases.py > $O/ncu_gemm_cases.log 2>&1
ls -la $O; cat $O/attn_time_s7_g4.log $O/attn_time_s5_g4.log; tail -9 $O/attn_fwd_s7.log; tail -n 4 $O/pytest_gpu.log; tail -n 4 $O/pytest_gpu_s7g4.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['loss'])"; done
