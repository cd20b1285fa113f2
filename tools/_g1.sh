mkdir -p gpurun_out
python -m pytest tests/test_gpu_upsample.py -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/u_pytest.log
timeout 300 python tools/bench_upsample.py --only-fused --steps 200 > gpurun_out/u_bench_cs.json 2> gpurun_out/u_bench_cs.err; echo "bench rc=$?"; cat gpurun_out/u_bench_cs.json; tail -5 gpurun_out/u_bench_cs.err
timeout 300 python tools/bench_upsample.py --only-fused --steps 200 --shape pascal > gpurun_out/u_bench_pascal.json 2> gpurun_out/u_bench_pascal.err; cat gpurun_out/u_bench_pascal.json
