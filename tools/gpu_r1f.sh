mkdir -p gpurun_out
for wl in pascal; do
timeout 300 python bench.py --workload $wl --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1f_bench_$wl.json 2> gpurun_out/r1f_bench_$wl.err; python -c "import json;d=json.load(open('gpurun_out/r1f_bench_$wl.json'));print('$wl tma',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
done
timeout 300 python bench.py --workload pascal --mode probs --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1f_bench_pascal_probs.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r1f_bench_pascal_probs.json'));print('pascal probs',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
timeout 300 python bench.py --workload pascal --mode votes --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1f_bench_pascal_votes.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r1f_bench_pascal_votes.json'));print('pascal votes',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
DAS_MC_TMA=0 timeout 300 python bench.py --workload pascal --mode votes --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1f_bench_pascal_votes_ldg.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r1f_bench_pascal_votes_ldg.json'));print('pascal votes ldg',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
timeout 200 python -m pytest tests/test_gpu_mc.py -m gpu -x -q 2>&1 | tail -2
