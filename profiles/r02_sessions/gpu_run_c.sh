# GPU session C (round 2, `gpurun --gpus 2`): multi-rank numerics and the overlapped all-reduce buckets on real NCCL.
set -x
O=gpurun_out/r2c
mkdir -p $O
nvidia-smi -L > $O/smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py $O/ddp_check.json > $O/ddp_check.log 2>&1; echo "rc=$?" >> $O/ddp_check.log
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q > $O/pytest_ddp.log 2>&1; echo "rc=$?" >> $O/pytest_ddp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg3 > $O/bench_2gpu_cfg3.json 2> $O/bench_2gpu_cfg3.err; echo "rc=$?" >> $O/bench_2gpu_cfg3.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_1gpu.json 2> $O/bench_1gpu.err
tail -n 5 $O/ddp_check.log; tail -n 3 $O/pytest_ddp.log; tail -c 400 $O/bench_2gpu.json; tail -c 300 $O/bench_2gpu.err
