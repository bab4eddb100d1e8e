# GPU session Z (round 2, `gpurun --gpus 4`): e2e at 4 ranks with / without NUMA-local pinned staging buffers.
set -x
O=gpurun_out/r2z
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
Q="--gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity"
B200_BENCH_NUMA_BIND=0 timeout -k 10 120 $TR --master-port 29621 bench.py $Q > $O/bench4_nobind.json 2> $O/bench4_nobind.err
B200_BENCH_NUMA_BIND=1 timeout -k 10 120 $TR --master-port 29622 bench.py $Q > $O/bench4_bind.json 2> $O/bench4_bind.err
for f in $O/bench4_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'))"; done; head -12 $O/topo.txt
