mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r1d_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; python -c "import json;d=json.load(open('gpurun_out/r1d_bench.json'));print(d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
timeout 300 python tools/bench_kcenter.py > gpurun_out/k4_bench3.json 2> gpurun_out/k4_bench3.err; cat gpurun_out/k4_bench3.json; tail -3 gpurun_out/k4_bench3.err
DAS_KC_CLUSTER=0 timeout 300 python tools/bench_kcenter.py > gpurun_out/k4_bench3_chain.json 2> gpurun_out/k4_bench3_chain.err; cat gpurun_out/k4_bench3_chain.json
