mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo pytest rc=$?
python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo smoke rc=$?
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err; echo ref rc=$?
python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_launch.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:mc_score_kernel -s 3 -c 2 -o gpurun_out/r1_prof_fused python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1; echo full rc=$?
