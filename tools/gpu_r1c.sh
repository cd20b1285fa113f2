mkdir -p gpurun_out
python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; echo bench rc=$?
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1c_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1c_ncu_launch.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_tma -c 2 -o gpurun_out/r1c_prof_tma python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1c_ncu_full.log 2>&1; echo full rc=$?
