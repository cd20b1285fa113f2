timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo bench rc=$?
python tools/bench_upsample.py > gpurun_out/g_up_cs.json 2>gpurun_out/g_up.err
python tools/bench_upsample.py --shape pascal > gpurun_out/g_up_pascal.json 2>>gpurun_out/g_up.err
tail -c 600 gpurun_out/g_up_cs.json; tail -c 600 gpurun_out/g_up_pascal.json
