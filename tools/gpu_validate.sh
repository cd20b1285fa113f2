#!/bin/bash
# One-GPU validation recipe (run on the B200 box:  gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh').
# Everything lands in gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v_pytest.log
python __graft_entry__.py smoke > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err; echo "reference arm rc=$?"
python bench.py --workload pascal --mode probs --no-e2e --no-cpu-baseline > gpurun_out/v_bench_pascal_probs.json 2>&1
python tools/bench_kcenter.py > gpurun_out/v_k4.json 2> gpurun_out/v_k4.err
python tools/bench_region.py > gpurun_out/v_region.json 2> gpurun_out/v_region.err
python tools/bench_upsample.py > gpurun_out/v_upsample_cs.json 2> gpurun_out/v_upsample.err
python tools/bench_upsample.py --shape pascal > gpurun_out/v_upsample_pascal.json 2>> gpurun_out/v_upsample.err
python tools/probe_selector.py --profile > gpurun_out/v_selector_profile.log 2>&1
python tools/probe_streaming.py --G 1 4 > gpurun_out/v_streaming.log 2>&1
# (profiles of round 2: tools/gpu_profile_r2.sh, tools/ncu_streaming.sh; fabric: tools/probe_h2d.py under torchrun)
# launch list of the timed region (bench.py brackets it with cudaProfilerStart/Stop), then the top kernel in full
CMD="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-upsample-variant --no-sweep --no-configs"
$CMD > gpurun_out/v_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches.csv $CMD > gpurun_out/v_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_tma -c 2 -o gpurun_out/v_prof_tma $CMD > gpurun_out/v_ncu_full.log 2>&1
UCMD="python tools/bench_upsample.py --only-fused --steps 10 --warmup 3"
$UCMD > gpurun_out/v_up_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mc_score_up -c 2 -o gpurun_out/v_prof_up $UCMD > gpurun_out/v_ncu_up.log 2>&1
echo done
