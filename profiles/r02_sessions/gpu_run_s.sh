# GPU session S (round 2): source-level ncu of b200_ln_gemm (where does the panel time go?).
set -x
O=gpurun_out/r2s
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_ln" -c 4 -o $O/ln_gemm python tools/ln_gemm_cases.py > $O/ncu_ln_gemm.log 2>&1
tail -3 $O/ncu_ln_gemm.log
