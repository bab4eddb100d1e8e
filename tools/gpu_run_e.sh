# GPU session E (round 2, `gpurun --gpus 2`): Sinkhorn-Knopp all-reduces captured into the step graph (B200_GRAPH_NCCL=1).
set -x
O=gpurun_out/r2e
mkdir -p $O
export B200_GRAPH_NCCL=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/ddp_check.py $O/ddp_check_graph_nccl.json > $O/ddp_check_graph_nccl.log 2>&1; echo "rc=$?" >> $O/ddp_check_graph_nccl.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_2gpu_cfg3_graph_nccl.json 2> $O/bench_2gpu_cfg3_graph_nccl.err; echo "rc=$?" >> $O/bench_2gpu_cfg3_graph_nccl.err
export B200_GRAPH_NCCL=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_2gpu_cfg3_eager.json 2> $O/bench_2gpu_cfg3_eager.err; echo "rc=$?" >> $O/bench_2gpu_cfg3_eager.err
timeout 300 python bench.py --steps 10 --warmup 3 --config cfg3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_1gpu_cfg3.json 2> $O/bench_1gpu_cfg3.err
tail -n 4 $O/ddp_check_graph_nccl.log; tail -c 300 $O/bench_2gpu_cfg3_graph_nccl.err; head -c 300 $O/bench_2gpu_cfg3_graph_nccl.json; head -c 300 $O/bench_2gpu_cfg3_eager.json
