mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1g_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r1g_pytest.log
timeout 600 python tools/bench_region.py > gpurun_out/r1g_region.json 2> gpurun_out/r1g_region.err; cat gpurun_out/r1g_region.json; tail -3 gpurun_out/r1g_region.err
