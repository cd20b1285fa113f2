#!/usr/bin/env python
"""Streaming form of K1 (das_mc_accumulate, G passes per launch, running fp32 state between launches) with and without
the L2-persisting window on the state: ms per batch and the algorithmic fraction of the HBM peak, for B in {1, 2, 8}.
Run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:mc_accumulate` to get the DRAM traffic.

    python tools/probe_streaming.py [--B 2] [--G 1] [--persist 1] [--reps 10] [--once]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, nargs="+", default=[1, 2, 8])
    ap.add_argument("--G", type=int, nargs="+", default=[1])
    ap.add_argument("--persist", type=int, nargs="+", default=[0, 1])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--once", action="store_true", help="one warm batch + one profiled batch (for ncu)")
    a = ap.parse_args()
    from deep_active_semantic_segmentation_b200 import _lib, ops, synth

    H, W, C, T = 512, 1024, 19, 20
    dev = torch.device("cuda", 0)
    passes, labels = synth.device_pass_logits(1, 0, max(a.B), T, C, H, W, dev)
    peak = 6548.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    print(json.dumps(_lib.l2_info(dev)))
    for B in a.B:
        sub = [p[:B] for p in passes]
        lab = labels[:B]
        scores = torch.zeros((B, _lib.N_SCORES), dtype=torch.float32, device=dev)
        for G in a.G:
            for persist in a.persist:
                _lib.set_option("mc_l2_persist", persist, dev)
                st = ops.MCState(B, C, H, W, T, votes=True, probs=True, device=dev, single_shot=(G >= T))
                groups = [sub[t0:t0 + G] for t0 in range(0, T, G)]

                def batch():
                    st.reset()
                    for g in groups[:-1]:
                        st.accumulate(g)
                    st.score(groups[-1], lab, maps=(), scores_out=scores)

                batch()
                torch.cuda.synchronize()
                if a.once:
                    torch.cuda.profiler.start()
                    batch()
                    torch.cuda.synchronize()
                    torch.cuda.profiler.stop()
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    batch()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.reps
                alg = T * B * C * H * W * 4
                print(json.dumps({"B": B, "G": G, "l2_persist": persist, "ms_per_batch": round(ms, 4),
                                  "images_per_s": round(B / ms * 1e3, 1), "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4),
                                  "state_MB": round(B * (C + 1) * H * W * 4 / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
