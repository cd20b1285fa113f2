mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r1e_pytest.log
for wl in pascal cityscapes; do
timeout 300 python bench.py --workload $wl --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1e_bench_$wl.json 2> gpurun_out/r1e_bench_$wl.err; python -c "import json;d=json.load(open('gpurun_out/r1e_bench_$wl.json'));print('$wl tma',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
done
DAS_MC_TMA=0 timeout 300 python bench.py --workload pascal --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1e_bench_pascal_ldg.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r1e_bench_pascal_ldg.json'));print('pascal ldg',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
timeout 300 python bench.py --workload pascal --mode probs --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1e_bench_pascal_probs.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/r1e_bench_pascal_probs.json'));print('pascal probs',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
