# GPU session B (round 2): measurement.  bench lines (cfg2 headline + cfg3/cfg5/cfg4), ncu launch list of one step,
# `ncu --set full` captures of the loss / optimizer / attention / GEMM kernels.  Outputs -> gpurun_out/r2b/
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
timeout 600 python bench.py --steps 20 --warmup 5 --gemm-profile $O/gemm_shapes.csv --no-cpu-baseline --no-gpu-baseline --no-parity > $O/bench_cfg2_b.json 2>> $O/bench_cfg2.err
# launch list of ONE eagerly launched step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python tools/profile_step.py 64 > $O/profile_step.log 2>&1
# full captures: loss path, optimizer sweep, attention, one of each GEMM flavour (few launches each: ~40 replays per launch)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"dino_ce|row_lse|adamw_ema|sumsq" -c 8 -o $O/loss_optim python tools/profile_step.py 64 > $O/ncu_loss.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_" -c 6 -o $O/attn python tools/profile_step.py 64 > $O/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"gemm_tcgen05" -s 20 -c 12 -o $O/gemm python tools/profile_step.py 64 > $O/ncu_gemm.log 2>&1
for c in cfg3 cfg5 cfg4; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err; echo "rc=$?" >> $O/bench_$c.err
done
ls -la $O; tail -c 400 $O/bench_cfg2.err; for c in cfg3 cfg5 cfg4; do tail -c 300 $O/bench_$c.err; done
