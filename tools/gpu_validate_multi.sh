#!/bin/bash
# Multi-GPU validation (gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_validate_multi.sh N'): sharded-selector parity on
# every rank against the un-sharded goldens, then the bench lines for N ranks: config 2 (bench.py), config 4 shape
# (bench.py --workload pascal --mode probs) and config 3 through the selector API (tools/bench_region_dist.py).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29517 tools/dist_check.py > gpurun_out/vm_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; tail -2 gpurun_out/vm_dist_check_n$N.log
$TR --master-port 29518 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/vm_bench_n$N.json 2> gpurun_out/vm_bench_n$N.err; echo "bench rc=$?"; wc -l gpurun_out/vm_bench_n$N.json
$TR --master-port 29519 bench.py --gpus $N --workload pascal --mode probs --steps 200 --warmup 5 --no-e2e > gpurun_out/vm_bench_pascal_probs_n$N.json 2> gpurun_out/vm_bench_pascal_n$N.err; echo "pascal rc=$?"
$TR --master-port 29520 tools/bench_region_dist.py > gpurun_out/vm_region_selector_n$N.json 2> gpurun_out/vm_region_n$N.err; echo "region rc=$?"
$TR --master-port 29521 tools/bench_region_dist.py --lowres > gpurun_out/vm_region_selector_lowres_n$N.json 2>> gpurun_out/vm_region_n$N.err; echo "region lowres rc=$?"
