# GPU session Q (round 2): LayerNorm-prologue GEMM (b200_ln_gemm): numerics, timing against LN + GEMM, step with B200_LN_GEMM=1.
set -x
O=gpurun_out/r2q
mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "ln_gemm" > $O/pytest_ln_gemm.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ln_gemm.log
timeout 300 python tools/ln_gemm_bench.py > $O/ln_gemm_bench.log 2>&1; echo "rc=$?" >> $O/ln_gemm_bench.log
B200_LN_GEMM=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -x > $O/pytest_parity_lngemm.log 2>&1; echo "pytest rc=$?" >> $O/pytest_parity_lngemm.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-e2e"
timeout 300 python bench.py $Q > $O/bench_default.json 2>> $O/bench_ab.err
B200_LN_GEMM=1 timeout 300 python bench.py $Q > $O/bench_lngemm.json 2>> $O/bench_ab.err
cat $O/ln_gemm_bench.log; tail -n 12 $O/pytest_ln_gemm.log; tail -n 3 $O/pytest_parity_lngemm.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['loss'], d['parity'] and d['parity']['loss_delta_vs_oracle'])"; done
