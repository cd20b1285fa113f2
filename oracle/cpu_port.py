"""TEST / BASELINE INFRASTRUCTURE - multi-threaded torch-CPU port of the reference's scoring op
sequence, used only as the timed CPU baseline (bench.py `cpu_baseline` and `--impl reference`).

The reference is pure Python and cannot travel to the GPU box, so its CPU path is restated here
with the same ATen ops it issues, in the same order (kind = "port"):
  votes      torch.argmax(logits, dim=1) stored as float32                  mc_dropout.py:37-40
  histogram  per class: sum(outputs == c, dim=0, dtype=f32) / T, -p*log2(p+1e-12)   mc_dropout.py:46-48
  softmax    nn.Softmax2d, per class -p*log2(p+1e-12)                        ceal.py:111-118
  image mean torch.mean                                                      mc_dropout.py:189
The composed scores (predictive entropy of the MC mean, BALD) reuse those primitives.
tests/test_oracle_vs_golden.py checks this port against oracle/restate.py.
"""
from __future__ import annotations

import torch


def score_batch(pass_logits, labels, C: int, want_probs: bool = True):
    """pass_logits: list of T float32 tensors [B,C,H,W] (CPU for the CPU baseline; CUDA tensors give the
    "same-box PyTorch eager" comparator); labels [B,H,W] float32.
    -> dict of float32 [B] image scores (vote_entropy, and with want_probs pred_entropy / bald /
    expected_entropy / confidence / margin)."""
    T = len(pass_logits)
    B, _, H, W = pass_logits[0].shape
    dev = pass_logits[0].device
    outputs = torch.empty(B, T, H, W, dtype=torch.float32, device=dev)
    softmax = torch.nn.Softmax2d()
    if want_probs:
        p_sum = torch.zeros(B, C, H, W, dtype=torch.float32, device=dev)
        e_sum = torch.zeros(B, H, W, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for t, x in enumerate(pass_logits):
            outputs[:, t] = torch.argmax(x, dim=1)
            if want_probs:
                p = softmax(x)
                p_sum += p
                e = torch.zeros(B, H, W, dtype=torch.float32, device=dev)
                for c in range(C):
                    e = e - p[:, c] * torch.log2(p[:, c] + 1e-12)
                e_sum += e
        res = {k: [] for k in (("vote_entropy", "pred_entropy", "bald", "expected_entropy", "confidence", "margin")
                               if want_probs else ("vote_entropy",))}
        for i in range(B):
            mask = (labels[i] < 0) | (labels[i] >= C)
            ve = torch.zeros(H, W, dtype=torch.float32, device=dev)
            for c in range(C):
                p = torch.sum(outputs[i] == c, dim=0, dtype=torch.float32) / T
                ve = ve - p * torch.log2(p + 1e-12)
            ve[mask] = 0
            res["vote_entropy"].append(torch.mean(ve))
            if want_probs:
                pb = p_sum[i] / T
                pe = torch.zeros(H, W, dtype=torch.float32, device=dev)
                for c in range(C):
                    pe = pe - pb[c] * torch.log2(pb[c] + 1e-12)
                ee = e_sum[i] / T
                pe[mask] = 0
                ee[mask] = 0
                top2 = torch.topk(pb, 2, dim=0).values
                conf, marg = top2[0].clone(), top2[0] - top2[1]
                conf[mask] = 1
                marg[mask] = 1
                res["pred_entropy"].append(torch.mean(pe))
                res["expected_entropy"].append(torch.mean(ee))
                res["bald"].append(torch.mean(pe - ee))
                res["confidence"].append(torch.mean(conf))
                res["margin"].append(torch.mean(marg))
    return {k: torch.stack(v) for k, v in res.items()}
