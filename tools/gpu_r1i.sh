mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r1i_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r1i_smoke.log
python tools/bench_kcenter.py --reps 2 > gpurun_out/r1i_k4_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kcenter_cluster -c 1 -o gpurun_out/r1i_prof_cluster python tools/bench_kcenter.py --reps 2 > gpurun_out/r1i_ncu_cluster.log 2>&1; echo ncu rc=$?
