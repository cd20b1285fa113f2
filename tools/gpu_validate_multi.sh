#!/bin/bash
# Multi-GPU validation (gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_validate_multi.sh N'): sharded-selector parity on
# every rank against the un-sharded goldens, then the bench line for N ranks.
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/vm_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; tail -2 gpurun_out/vm_dist_check_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/vm_bench_n$N.json 2> gpurun_out/vm_bench_n$N.err; echo "bench rc=$?"; wc -l gpurun_out/vm_bench_n$N.json
