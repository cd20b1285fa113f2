mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tma_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/tma_pytest.log
for ctas in 3 2 4; do
  DAS_MC_TMA_CTAS=$ctas timeout 300 python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/tma_bench_c$ctas.json 2> gpurun_out/tma_bench_c$ctas.err; echo "ctas=$ctas rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/tma_bench_c$ctas.json'));print(d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
done
DAS_MC_TMA=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/tma_bench_off.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/tma_bench_off.json'));print('ldg',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"
for mode in votes probs; do timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --mode $mode > gpurun_out/tma_bench_$mode.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/tma_bench_$mode.json'));print('$mode',d['value'],d['roofline']['frac'],d['roofline']['avg_launch_ms'],d['clocks'])"; done
timeout 300 python tools/bench_kcenter.py > gpurun_out/k4_bench2.json 2> gpurun_out/k4_bench2.err; cat gpurun_out/k4_bench2.json
