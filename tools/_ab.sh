for v in 0 220; do
  echo "== variant $v"
  DAS_MC_UP_WARPS=$v python tools/bench_upsample.py --only-fused --steps 200 --warmup 10 2>&1 | tail -1 | cut -c1-180
  DAS_MC_UP_WARPS=$v python tools/bench_upsample.py --only-fused --steps 200 --warmup 10 --shape pascal 2>&1 | tail -1 | cut -c1-180
done
timeout 300 python -m pytest tests/test_gpu_upsample.py -q -m gpu -x 2>&1 | tail -3
