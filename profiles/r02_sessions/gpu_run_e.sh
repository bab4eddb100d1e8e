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
# where does the 2-GPU step lose its 3 %?  cfg2: default | cut at depth/4 | no all-reduce at all (diagnostic ceiling) | fewer NCCL channels
Q="--gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29561 bench.py $Q > $O/bench2_default.json 2>> $O/bench2.err
B200_DDP_SPLIT_FRAC=0.25 timeout 300 $TR --master-port 29562 bench.py $Q > $O/bench2_split025.json 2>> $O/bench2.err
B200_DDP_SPLIT_FRAC=0.75 timeout 300 $TR --master-port 29563 bench.py $Q > $O/bench2_split075.json 2>> $O/bench2.err
B200_BENCH_DIAG_NO_ALLREDUCE=1 timeout 300 $TR --master-port 29564 bench.py $Q > $O/bench2_diag_noallreduce.json 2>> $O/bench2.err
NCCL_MAX_NCHANNELS=8 timeout 300 $TR --master-port 29565 bench.py $Q > $O/bench2_nch8.json 2>> $O/bench2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e > $O/bench1_default.json 2>> $O/bench2.err
for f in $O/bench2_*.json $O/bench1_default.json; do echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['host_ms_per_step'])"; done
tail -n 4 $O/ddp_check_graph_nccl.log; tail -c 300 $O/bench_2gpu_cfg3_graph_nccl.err; head -c 300 $O/bench_2gpu_cfg3_graph_nccl.json; head -c 300 $O/bench_2gpu_cfg3_eager.json
