mkdir -p gpurun_out
CMD="python tools/bench_upsample.py --only-fused --steps 10 --warmup 3"
$CMD > gpurun_out/u_plain.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_up -c 2 -o gpurun_out/u_prof_up $CMD > gpurun_out/u_ncu_full.log 2>&1
echo ncu rc=$?
tail -3 gpurun_out/u_ncu_full.log
