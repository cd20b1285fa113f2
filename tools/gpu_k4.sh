mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_region_topk_kcenter.py -m gpu -x -q -k "gram or filtered or sharded_steps or kcenter" > gpurun_out/k4_pytest.log 2>&1; echo pytest rc=$?
tail -30 gpurun_out/k4_pytest.log
timeout 300 python tools/bench_kcenter.py --cpu > gpurun_out/k4_bench.json 2> gpurun_out/k4_bench.err; echo bench rc=$?
cat gpurun_out/k4_bench.json; tail -5 gpurun_out/k4_bench.err
