mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1l_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r1l_pytest.log
python bench.py --steps 400 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; python -c "import json;d=json.load(open('gpurun_out/r1l_bench.json'));print(d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
python bench.py --workload pascal --steps 400 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1l_bench_p.json 2> gpurun_out/r1l_bench_p.err; python -c "import json;d=json.load(open('gpurun_out/r1l_bench_p.json'));print(d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
