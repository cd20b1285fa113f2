mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1j_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r1j_pytest.log
python bench.py > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; echo bench rc=$?; cat gpurun_out/r1j_bench.json; tail -3 gpurun_out/r1j_bench.err
