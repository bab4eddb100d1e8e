# GPU session O (round 2, `gpurun --gpus 2`): what grew the 2-GPU gap (0.38 -> 0.85 ms)?  same box: 1 GPU | 2 GPUs default |
# 8-warp backward | warp-level forward | no all-reduce (diagnostic)
set -x
O=gpurun_out/r2o
mkdir -p $O
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 300 python bench.py $Q > $O/bench1.json 2>> $O/err.log
timeout -k 10 300 $TR --master-port 29591 bench.py --gpus 2 $Q > $O/bench2_default.json 2>> $O/err.log
B200_ATTN_BWD_GROUPS=2 timeout -k 10 300 $TR --master-port 29592 bench.py --gpus 2 $Q > $O/bench2_bwdg2.json 2>> $O/err.log
B200_TC_ATTN_FWD=0 timeout -k 10 300 $TR --master-port 29593 bench.py --gpus 2 $Q > $O/bench2_warpfwd.json 2>> $O/err.log
B200_BENCH_DIAG_NO_ALLREDUCE=1 timeout -k 10 300 $TR --master-port 29594 bench.py --gpus 2 $Q > $O/bench2_diag_noallreduce.json 2>> $O/err.log
timeout -k 10 300 $TR --master-port 29595 bench.py --gpus 2 $Q > $O/bench2_default_b.json 2>> $O/err.log
for f in $O/bench*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['host_ms_per_step'])"; done
