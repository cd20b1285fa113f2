#!/usr/bin/env python
"""Class-count sweep of the single-shot resident-logits kernel (das_mc_accumulate_finalize, TMA ring) on one B200:
algorithmic GB/s per class count C at 512 x 1024, T = 20, B = 4 (every C is its own template instantiation).

    python tools/probe_classes.py [C ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from deep_active_semantic_segmentation_b200 import _lib, ops, synth

    classes = [int(v) for v in sys.argv[1:]] or [19, 21, 22, 24, 28, 32]
    peak = 6548.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    dev = torch.device("cuda", 0)
    B, T, H, W = 4, 20, 512, 1024
    for C in classes:
        logits, labels = synth.device_pass_logits(synth.DEFAULT_SEED, 0, B, T, C, H, W, dev, block=32)
        st = ops.MCState(B, C, H, W, T, device=dev, single_shot=True)
        scores = torch.zeros((B, _lib.N_SCORES), dtype=torch.float32, device=dev)

        def step():
            st.reset()
            st.score(logits, labels, maps=(), scores_out=scores)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gbs = T * B * C * H * W * 4 / (ms * 1e-3) / 1e9
        print(json.dumps({"C": C, "ms_per_batch": round(ms, 4), "algorithmic_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}), flush=True)
        del logits, labels, st


if __name__ == "__main__":
    main()
