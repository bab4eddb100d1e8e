# GPU session G (round 2, `gpurun --gpus 2`): Sinkhorn all-reduces captured in the step graph -- clean teardown after
# release_graphs() (session E: numerics fine, but destroy_process_group hung while the graphs were alive).
set -x
O=gpurun_out/r2g
mkdir -p $O
export B200_GRAPH_NCCL=1
date +%s > $O/t0
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/ddp_check.py $O/ddp_check_graph_nccl.json > $O/ddp_check_graph_nccl.log 2>&1; echo "rc=$?" >> $O/ddp_check_graph_nccl.log
date +%s > $O/t1
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_2gpu_cfg3_graph_nccl.json 2> $O/bench_2gpu_cfg3_graph_nccl.err; echo "rc=$?" >> $O/bench_2gpu_cfg3_graph_nccl.err
date +%s > $O/t2
tail -n 3 $O/ddp_check_graph_nccl.log | cut -c1-300; tail -n 2 $O/bench_2gpu_cfg3_graph_nccl.err; head -c 300 $O/bench_2gpu_cfg3_graph_nccl.json; cat $O/t0 $O/t1 $O/t2
