mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29519 tools/bench_region_dist.py > gpurun_out/m_region_n$N.json 2> gpurun_out/m_region_n$N.err; echo "region rc=$?"; cat gpurun_out/m_region_n$N.json
$TR --master-port 29520 tools/bench_region_dist.py --lowres > gpurun_out/m_region_lowres_n$N.json 2> gpurun_out/m_region_lowres_n$N.err; echo "region lowres rc=$?"; cat gpurun_out/m_region_lowres_n$N.json
$TR --master-port 29521 bench.py --gpus $N --workload pascal --mode probs --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/m_pascal_n$N.json 2> gpurun_out/m_pascal_n$N.err; echo "pascal rc=$?"; cut -c1-200 gpurun_out/m_pascal_n$N.json
