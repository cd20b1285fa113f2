#!/usr/bin/env python
"""BASELINE config 3 (region-based vote-entropy scoring, 128x128 regions, Cityscapes-shaped pool) on one B200.

Times, with CUDA events, the stages of ActiveSelectionMCDropout.create_region_maps for a shard of N images:
  score   T=20 passes -> vote-entropy map (fused vote kernel)              per batch of 8
  region  suppress labelled rects + RxR box sums + pool min/max           per batch of 8
  tail    min-max normalise + per-image NMS sequences + K3 global order + stop rule   once per pool
    python tools/bench_region.py [--N 256] [--R 128] [--k 125]
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
    ap.add_argument("--N", type=int, default=256)
    ap.add_argument("--R", type=int, default=128)
    ap.add_argument("--k", type=int, default=125)
    ap.add_argument("--H", type=int, default=512)
    ap.add_argument("--W", type=int, default=1024)
    a = ap.parse_args()
    import math
    import torch
    from deep_active_semantic_segmentation_b200 import dist, ops, synth

    H, W, C, T, B, R = a.H, a.W, 19, 20, 8, a.R
    dev = torch.device("cuda", 0)
    passes, labels = synth.device_pass_logits(synth.DEFAULT_SEED, 0, B, T, C, H, W, dev)
    H2, W2 = H - R + 1, W - R + 1
    nb = a.N // B
    score_maps = torch.empty((nb * B, H2, W2), dtype=torch.float32, device=dev)
    rects = [(b, 64 * (b % 3), 128 * (b % 5), 128, 128) for b in range(B) if b % 2 == 0]

    def ev():
        return torch.cuda.Event(enable_timing=True)

    st = ops.MCState(B, C, H, W, T, votes=True, probs=False, device=dev, single_shot=True)
    t_score = t_region = 0.0
    mm = ops.new_minmax(dev)
    for it in range(nb + 2):
        e0, e1, e2 = ev(), ev(), ev()
        i = max(it - 2, 0)                      # two warm-up batches
        st.reset()
        e0.record()
        # per-image jitter so that the pool is not 32 copies of one batch
        out = st.score(passes, labels, maps=("vote_entropy",), scores=False)
        e1.record()
        maps = out["vote_entropy"]
        maps.mul_(1.0 + 0.01 * ((i * 7) % 13))
        e1b = ev(); e1b.record()
        ops.suppress_rects(maps, rects)
        ops.box_sum(maps, R, mm, out=score_maps[i * B:(i + 1) * B])
        e2.record()
        torch.cuda.synchronize()
        if it >= 2:
            t_score += e0.elapsed_time(e1)
            t_region += e1b.elapsed_time(e2)

    num_requested = (a.k * H * W) / (R * R)
    kmax = max(1, min(math.ceil(num_requested), ops.nms_pick_bound(H2, W2, R)))
    from deep_active_semantic_segmentation_b200.active_selection import base
    keep = score_maps.clone()
    e0, e1 = ev(), ev()
    e0.record()
    ops.minmax_normalise(score_maps, mm)
    t0 = time.perf_counter()
    regions, count = base.global_nms(score_maps, 0, nb * B, R, num_requested, kmax)   # NMS kernel + K3 + D2H of the winners
    e1.record()
    torch.cuda.synchronize()
    t_tail_wall = (time.perf_counter() - t0) * 1e3
    t_dev_tail = e0.elapsed_time(e1)
    # the host k-way merge it replaces, for comparison
    score_maps.copy_(keep)
    ops.minmax_normalise(score_maps, mm)
    cs, rc, cnt = ops.nms_sequences(score_maps, R, kmax, 0.01)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seqs = dist.sequences_from_device(cs, rc, cnt)
    regions_h, count_h = dist.merge_nms_sequences(seqs, R, num_requested, H2, W2)
    t_host_tail = (time.perf_counter() - t0) * 1e3
    assert (regions_h, count_h) == (regions, count), "device merge != host k-way merge"

    n = nb * B
    logits_bytes = T * C * H * W * 4
    line = {
        "workload": f"region_vote_entropy_{H}x{W}_c{C}_t{T}_R{R}_k{a.k}", "images": n, "batch": B,
        "score_ms_per_image": round(t_score / n, 4), "score_GBps": round(logits_bytes * n / (t_score * 1e-3) / 1e9, 1),
        "region_ms_per_image": round(t_region / n, 4),
        "region_bytes_per_image": H * W * 4 + H2 * W2 * 4,
        "tail_ms": round(t_dev_tail, 3), "tail_ms_per_image": round(t_dev_tail / n, 4), "tail_wall_ms": round(t_tail_wall, 3),
        "replaced_host_merge_ms": round(t_host_tail, 3), "kmax_per_image": kmax, "picked": count,
        "images_per_s": round(n / ((t_score + t_region + t_dev_tail) * 1e-3), 1),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
