ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_up -c 1 -o gpurun_out/p_up_final_cs python tools/bench_upsample.py --only-fused --steps 2 --warmup 1 > gpurun_out/p_up_final_cs.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_up -c 1 -o gpurun_out/p_up_final_pascal python tools/bench_upsample.py --only-fused --steps 2 --warmup 1 --shape pascal > gpurun_out/p_up_final_pascal.log 2>&1
DAS_MC_UP_WARPS=15 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_up -c 1 -o gpurun_out/p_up_final_cs_pairs python tools/bench_upsample.py --only-fused --steps 2 --warmup 1 > gpurun_out/p_up_final_cs_pairs.log 2>&1
ls -la gpurun_out/p_up_final*
