#!/usr/bin/env python
"""BASELINE config 5 (core-set k-center greedy): N=10000, D=2048, L=50, K=500 on one B200.

Prints one JSON line: time of the tcgen05 bf16 distance GEMM (filter build), of the greedy loop with and
without the filter, the achieved tensor TFLOP/s against MEASURED_PEAKS.json, the share of rows the filter
sent to the exact float64 path, and (optionally) the reference's CPU time (oracle port of core_set.py:17-38).
    python tools/bench_kcenter.py [--N 10000 --D 2048 --L 50 --K 500] [--cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=10000)
    ap.add_argument("--D", type=int, default=2048)
    ap.add_argument("--L", type=int, default=50)
    ap.add_argument("--K", type=int, default=500)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from deep_active_semantic_segmentation_b200 import ops, synth

    feats_h = synth.coreset_features(11, a.N, a.D)
    feats = torch.from_numpy(feats_h).cuda()
    cen = list(range(a.L))

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts), sum(ts) / len(ts), out

    torch.cuda.profiler.start()
    t_build_min, t_build_avg, flt = timed(lambda: ops.KCenterFilter(feats), a.reps)
    t_f_min, t_f_avg, (p1, m1) = timed(lambda: ops.kcenter_greedy(feats, cen, a.K, flt), a.reps)
    torch.cuda.profiler.stop()
    flt = ops.KCenterFilter(feats)
    p1, m1 = ops.kcenter_greedy(feats, cen, a.K, flt)
    exact, screened = flt.stats()
    t_e_min, t_e_avg, (p0, m0) = timed(lambda: ops.kcenter_greedy(feats, cen, a.K), max(2, a.reps // 2))
    same = p0.cpu().tolist() == p1.cpu().tolist() and bool(torch.equal(m0, m1))

    Dp = -(-a.D // 64) * 64
    flops = 2.0 * a.N * a.N * Dp
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    line = {
        "workload": f"coreset_kcenter_N{a.N}_D{a.D}_L{a.L}_K{a.K}",
        "filter_build_ms": round(t_build_min, 4), "filter_build_avg_ms": round(t_build_avg, 4),
        "gemm_flops": flops, "gemm_tflops_incl_prepare": round(flops / (t_build_min * 1e-3) / 1e12, 1),
        "bf16_peak_tflops": peaks["bf16_tflops"],
        "tensor_frac_incl_prepare": round(flops / (t_build_min * 1e-3) / 1e12 / peaks["bf16_tflops"], 4),
        "greedy_filtered_ms": round(t_f_min, 4), "greedy_exact_ms": round(t_e_min, 4),
        "total_filtered_ms": round(t_build_min + t_f_min, 4),
        "speedup_vs_exact_fp64_gpu": round(t_e_min / (t_build_min + t_f_min), 2),
        "exact_rows": exact, "screened_rows": screened,
        "exact_fraction_of_steps": round((exact - 0) / max(screened, 1), 5),
        "filtered_equals_exact": same,
        "reference_algorithmic_flops": 2.0 * a.N * a.D * (a.L + a.K),
    }
    if a.cpu:
        from oracle import restate as R
        t0 = time.perf_counter()
        want, _ = R.kcenter_greedy(feats_h, cen, a.K)
        line["cpu_oracle_s"] = round(time.perf_counter() - t0, 3)
        line["cpu_cores"] = os.cpu_count()
        line["picks_equal_cpu_oracle"] = want == p1.cpu().tolist()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
