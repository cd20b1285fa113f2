mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/r1h_pytest.log
timeout 600 python tools/bench_region.py > gpurun_out/r1h_region.json 2> gpurun_out/r1h_region.err; cat gpurun_out/r1h_region.json; tail -3 gpurun_out/r1h_region.err
