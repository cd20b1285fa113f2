mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
for m in full probs; do python tools/bench_upsample.py --only-fused --steps 200 --mode $m 2>/dev/null | cut -c90-180; done
python tools/bench_upsample.py --only-fused --steps 200 --shape pascal 2>/dev/null | cut -c90-180
python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-upsample-variant 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['clocks'])"
python bench.py --steps 200 --no-e2e --no-cpu-baseline --no-upsample-variant --pass-group 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('G=1', d['value'], d['roofline']['frac'])"
