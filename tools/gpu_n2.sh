mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/n2_dist_check.log 2>&1; echo dist_check rc=$?; tail -5 gpurun_out/n2_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo bench rc=$?; tail -2 gpurun_out/n2_bench.json
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/n2_bench_n1.json 2>&1; tail -1 gpurun_out/n2_bench_n1.json | cut -c1-400
timeout 300 python tools/bench_kcenter.py > gpurun_out/k4_bench4.json 2> gpurun_out/k4_bench4.err; cat gpurun_out/k4_bench4.json
