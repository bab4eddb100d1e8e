# GPU session X (round 2, `gpurun --gpus 2`): final tree on 2 GPUs: DDP check, 2-GPU bench lines, reference arm under torchrun.
set -x
O=gpurun_out/r2x
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 300 $TR --master-port 29601 tools/ddp_check.py $O/ddp_check.json > $O/ddp_check.log 2>&1; echo "rc=$?" >> $O/ddp_check.log
timeout -k 10 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q > $O/pytest_ddp.log 2>&1; echo "rc=$?" >> $O/pytest_ddp.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_1gpu_cfg2.json 2> $O/bench_1gpu_cfg2.err
timeout -k 10 400 $TR --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $O/bench_2gpu_cfg2.json 2> $O/bench_2gpu_cfg2.err; echo "rc=$?" >> $O/bench_2gpu_cfg2.err
timeout -k 10 300 $TR --master-port 29603 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg3 --no-cpu-baseline --no-gpu-baseline > $O/bench_2gpu_cfg3.json 2> $O/bench_2gpu_cfg3.err; echo "rc=$?" >> $O/bench_2gpu_cfg3.err
tail -n 2 $O/ddp_check.log | cut -c1-300; tail -n 2 $O/pytest_ddp.log
for f in $O/bench_*.json; do python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'))"; done
