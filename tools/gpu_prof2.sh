mkdir -p gpurun_out
# launch list of the timed region only (cudaProfilerStart/Stop in bench.py)
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1b_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1b_ncu_launch.log 2>&1; echo launches rc=$?
# K4: launch list + full capture of the tcgen05 GEMM
python tools/bench_kcenter.py --reps 2 > gpurun_out/k4_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k4_launches.csv python tools/bench_kcenter.py --reps 2 > gpurun_out/k4_ncu_launch.log 2>&1; echo k4 launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:kc_dist_gemm -c 2 -o gpurun_out/k4_prof_gemm python tools/bench_kcenter.py --reps 2 > gpurun_out/k4_ncu_full.log 2>&1; echo k4 full rc=$?
