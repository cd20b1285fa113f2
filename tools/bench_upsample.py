#!/usr/bin/env python
"""MC scoring with the network's final bilinear upsample fused into the kernel, on one B200.

Per step (one batch of B images, T stochastic passes, the low-resolution decoder logits resident in HBM):
  fused     das_mc_upsample_accumulate_finalize on [B,C,h,w] x T                       (ONE launch)
  unfused   T x F.interpolate(low, (H,W), bilinear, align_corners=True)  +  das_mc_accumulate_finalize
            (what the reference model + the resident-logits kernel do: models/deeplab.py:59, then K1+K2)
  score     das_mc_accumulate_finalize alone on pre-interpolated logits (the BASELINE config 2 step)
Times are CUDA-event means over --steps launches after --warmup; one JSON line on stdout.
    python tools/bench_upsample.py [--shape cityscapes|pascal] [--steps 100] [--mode full|probs|votes]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="cityscapes", choices=["cityscapes", "pascal"])
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--mode", default="full", choices=["full", "probs", "votes"])
    ap.add_argument("--only-fused", action="store_true", help="profiling: launch nothing but the fused kernel")
    ap.add_argument("--classes", type=int, default=0, help="override the class count of the shape")
    a = ap.parse_args()
    import torch
    import torch.nn.functional as F
    from deep_active_semantic_segmentation_b200 import _lib, ops, synth

    if a.shape == "cityscapes":
        C, T, h, w, H, W = 19, 20, 128, 256, 512, 1024
    else:
        C, T, h, w, H, W = 21, 20, 129, 129, 513, 513
    if a.classes:
        C = a.classes
    B = a.batch
    dev = torch.device("cuda", 0)
    votes, probs = a.mode in ("full", "votes"), a.mode in ("full", "probs")
    low, lab_low = synth.device_pass_logits(synth.DEFAULT_SEED, 0, B, T, C, h, w, dev, block=8)
    labels = F.interpolate(lab_low[:, None], size=(H, W), mode="nearest")[:, 0].contiguous()
    st = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, device=dev, single_shot=True)
    scores = torch.zeros((B, _lib.N_SCORES), dtype=torch.float32, device=dev)

    def timed(fn, steps):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def fused():
        st.reset()
        st.score_upsampled(low, labels, maps=(), scores_out=scores)

    torch.cuda.profiler.start()
    ms_fused = timed(fused, a.steps)
    torch.cuda.profiler.stop()
    s_fused = scores.clone()
    out = {"shape": a.shape, "B": B, "T": T, "C": C, "h": h, "w": w, "H": H, "W": W, "mode": a.mode,
           "fused_ms_per_batch": round(ms_fused, 4), "fused_images_per_s": round(B / ms_fused * 1e3, 1),
           "lowres_bytes_per_batch": T * B * C * h * w * 4, "fullres_bytes_per_batch": T * B * C * H * W * 4}
    if not a.only_fused:
        up = [None] * T

        def interp():
            for t in range(T):
                up[t] = F.interpolate(low[t], size=(H, W), mode="bilinear", align_corners=True)  # the model's last op

        def score():
            st.reset()
            st.score(up, labels, maps=(), scores_out=scores)

        def unfused():
            interp()
            score()

        interp()
        ms_score = timed(score, a.steps)
        ms_unfused = timed(unfused, max(a.steps // 4, 5))
        torch.cuda.synchronize()
        diff = (scores - s_fused).abs().max().item()
        out.update({"score_only_ms_per_batch": round(ms_score, 4), "score_only_images_per_s": round(B / ms_score * 1e3, 1),
                    "interpolate_then_score_ms_per_batch": round(ms_unfused, 4),
                    "interpolate_then_score_images_per_s": round(B / ms_unfused * 1e3, 1),
                    "speedup_vs_interpolate_then_score": round(ms_unfused / ms_fused, 2),
                    "max_abs_score_difference": diff})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
