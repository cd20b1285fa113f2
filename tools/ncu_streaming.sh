#!/bin/bash
# DRAM traffic of the streaming form per batch (ncu), for profiles/r2_pass_group_traffic.json.
# --cache-control none + --replay-mode application: the point of the experiment is what stays in L2 BETWEEN launches,
# so ncu must neither flush the caches before a kernel nor replay a kernel in place.
mkdir -p gpurun_out
for cfg in "2 1 1" "2 1 0" "1 1 1" "1 1 0" "8 1 1" "2 4 1" "2 5 1" "2 10 1" "2 20 1" "8 4 1" "8 5 1" "8 10 1" "8 20 1"; do
  set -- $cfg
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --cache-control none --replay-mode application --csv \
      --log-file gpurun_out/ncu_stream_B$1_G$2_P$3.csv python tools/probe_streaming.py --B $1 --G $2 --persist $3 --once > /dev/null 2>&1
  echo "done $cfg"
done
