mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/n2_dist_check.log 2>&1; echo dist_check rc=$?; tail -3 gpurun_out/n2_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo bench rc=$?; head -c 200 gpurun_out/n2_bench.json; echo; wc -l gpurun_out/n2_bench.json
