mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/n${n}_bench.json 2> gpurun_out/n${n}_bench.err; echo "n=$n rc=$?"; tail -1 gpurun_out/n${n}_bench.json | cut -c1-330; tail -2 gpurun_out/n${n}_bench.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/n8_dist_check.log 2>&1; echo dist_check rc=$?; tail -2 gpurun_out/n8_dist_check.log
