# GPU session B (round 2): full -m gpu suite after the fixes, attention / loss A-B microbenches, bench A/B lines, ncu launch
# list + full captures.  Outputs -> gpurun_out/r2b/
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for mode in fwd bwd time; do timeout 240 python tools/attn_check.py $mode > $O/attn3_$mode.log 2>&1; echo "rc=$?" >> $O/attn3_$mode.log; done
B200_ATTN_FWD_SCHEDULE=1 timeout 240 python tools/attn_check.py time > $O/attn1_time.log 2>&1
for v in 0 4; do B200_CE_VARIANT=$v timeout 200 python tools/loss_bench.py >> $O/loss_bench.log 2>&1; done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?" >> $O/bench_cfg2.err
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-parity --no-e2e"
B200_TC_ATTN_FWD=1 timeout 300 python bench.py $Q > $O/bench_tcfwd.json 2>> $O/bench_ab.err
B200_TC_ATTN_FWD=1 B200_TC_ATTN_PACKED=1 timeout 300 python bench.py $Q > $O/bench_tcfwd_packed.json 2>> $O/bench_ab.err
B200_TC_ATTN_BWD=0 timeout 300 python bench.py $Q > $O/bench_warp_attn.json 2>> $O/bench_ab.err
B200_SUBSET_MIN_RATE=0 timeout 300 python bench.py $Q > $O/bench_subset_compact.json 2>> $O/bench_ab.err
for v in 0 1 2; do B200_CE_VARIANT=$v timeout 300 python bench.py $Q > $O/bench_ce$v.json 2>> $O/bench_ab.err; done
# launch list of ONE eagerly launched step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python tools/profile_step.py 64 > $O/profile_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"dino_ce|row_lse|adamw_ema|sumsq" -c 8 -o $O/loss_optim python tools/profile_step.py 64 > $O/ncu_loss.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_" -c 6 -o $O/attn python tools/profile_step.py 64 > $O/ncu_attn.log 2>&1
for c in cfg3 cfg5 cfg4; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err; echo "rc=$?" >> $O/bench_$c.err
done
ls -la $O; tail -n 12 $O/pytest_gpu.log; cat $O/attn3_time.log $O/attn1_time.log | grep time; tail -c 300 $O/bench_cfg2.err
