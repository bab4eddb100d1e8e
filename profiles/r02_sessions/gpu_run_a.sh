# GPU session A (round 2): tcgen05 attention bring-up checks, the -m gpu suite, one bench line.  Outputs -> gpurun_out/r2a/
set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
nproc >> gpurun_out/r2a/smi.txt
for mode in fwd bwd time; do
  timeout 240 python tools/attn_check.py $mode > gpurun_out/r2a/attn_$mode.log 2>&1; echo "rc=$?" >> gpurun_out/r2a/attn_$mode.log
done
B200_TEST_UNVALIDATED=1 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
B200_PDL=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q > gpurun_out/r2a/pytest_pdl.log 2>&1; echo "pdl pytest rc=$?" >> gpurun_out/r2a/pytest_pdl.log
B200_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity > gpurun_out/r2a/bench_pdl.json 2> gpurun_out/r2a/bench_pdl.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?" >> gpurun_out/r2a/bench.err
tail -n 4 gpurun_out/r2a/attn_fwd.log gpurun_out/r2a/attn_bwd.log gpurun_out/r2a/attn_time.log; tail -n 15 gpurun_out/r2a/pytest_gpu.log; tail -n 3 gpurun_out/r2a/pytest_pdl.log; tail -c 300 gpurun_out/r2a/bench.err
