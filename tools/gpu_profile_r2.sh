#!/bin/bash
# Round-2 profile captures on one B200 (gpurun --timeout 1500 -- 'bash tools/gpu_profile_r2.sh'); everything lands in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-upsample-variant --no-sweep --no-configs"
$CMD > gpurun_out/p_plain.json 2> gpurun_out/p_plain.err || { echo "plain bench failed"; exit 1; }
# launch list of the timed region (bench.py brackets it with cudaProfilerStart/Stop)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches.csv $CMD > gpurun_out/p_ncu_launch.log 2>&1
# the dominant kernel in full
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_tma -c 2 -o gpurun_out/p_prof_tma $CMD > gpurun_out/p_ncu_tma.log 2>&1
# K4: tcgen05 distance GEMM + cluster greedy loop
KCMD="python tools/bench_kcenter.py --reps 1"
$KCMD > gpurun_out/p_k4.json 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kcenter_cluster|kc_dist_gemm2" -c 2 -o gpurun_out/p_prof_k4 $KCMD > gpurun_out/p_ncu_k4.log 2>&1
# noise kernel + flat TMA kernel (config 4 shape)
PCMD="python bench.py --workload pascal --mode probs --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-upsample-variant --no-sweep --no-configs"
$PCMD > gpurun_out/p_pascal.json 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k regex:mc_score_tma -c 1 -o gpurun_out/p_prof_tma_flat $PCMD > gpurun_out/p_ncu_flat.log 2>&1
echo done
