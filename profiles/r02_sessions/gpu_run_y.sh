# GPU session Y (round 2, `gpurun --gpus 4`): 4-rank sanity of the committed tree: DDP check, bench line.
set -x
O=gpurun_out/r2y
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout -k 10 150 $TR --master-port 29611 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_4gpu_cfg2.json 2> $O/bench_4gpu_cfg2.err; echo "rc=$?" >> $O/bench_4gpu_cfg2.err
timeout -k 10 200 $TR --master-port 29612 tools/ddp_check.py $O/ddp_check.json > $O/ddp_check.log 2>&1; echo "rc=$?" >> $O/ddp_check.log
tail -n 2 $O/ddp_check.log | cut -c1-300; tail -n 2 $O/bench_4gpu_cfg2.err | cut -c1-200
python -c "import json; d=json.loads(open('$O/bench_4gpu_cfg2.json').read().strip().splitlines()[-1]); print(d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'))"
