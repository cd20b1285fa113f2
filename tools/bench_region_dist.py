#!/usr/bin/env python
"""BASELINE config 3 at N GPUs: region-based vote-entropy scoring (128 x 128 regions) with top-k region selection
through the selector API, the Cityscapes-shaped pool sharded by image over the ranks (weak scaling:
--images-per-rank images each).

    python tools/bench_region_dist.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
        tools/bench_region_dist.py

One call of ActiveSelectionMCDropout.create_region_maps(model, images, existing_regions, R, k) is timed
(barrier + synchronize on both sides, max over ranks): per batch T stochastic forwards (a replay model whose
logits are resident in HBM) -> fused vote kernel -> suppress labelled rectangles -> R x R box sums; then the
pool tail: min-max all-reduce, per-image NMS kernel, K3 over the candidate table, candidate all-gather, merge.
Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (init_dist / barrier / max_over_ranks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images-per-rank", type=int, default=256)
    ap.add_argument("--R", type=int, default=128)
    ap.add_argument("--k", type=int, default=125)
    ap.add_argument("--lowres", action="store_true", help="the model returns its low-resolution decoder logits")
    a = ap.parse_args()
    import torch
    from deep_active_semantic_segmentation_b200 import _lib, constants, synth
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCDropout, base

    H, W, C, T, B = 512, 1024, 19, 20, 8
    world, rank, local = bench.init_dist(int(os.environ.get("WORLD_SIZE", "1")))
    dev = torch.device("cuda", local)
    n_total = world * a.images_per_rank
    h, w = (H // 4, W // 4) if a.lowres else (H, W)
    passes, lab = synth.device_pass_logits(synth.DEFAULT_SEED + rank, rank * a.images_per_rank, B, T, C, h, w, dev,
                                           block=8 if a.lowres else 32)
    labels = lab if not a.lowres else torch.nn.functional.interpolate(lab[:, None], size=(H, W), mode="nearest")[:, 0].contiguous()
    image = torch.zeros(3, H, W).pin_memory()
    labels_h = labels.cpu().pin_memory()

    class ResidentDataset(torch.utils.data.Dataset):   # images / labels stand-ins; the logits never leave the device
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            return {"image": image, "label": labels_h[int(self.paths[i]) % B]}

    class ResidentReplayModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.t = 0

        def forward(self, x):
            out = passes[self.t % T]
            self.t += 1
            return out[:x.shape[0]]

    old_ds, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
    base.paths_dataset.PathsDataset, constants.MC_STEPS = ResidentDataset, T
    try:
        sel = ActiveSelectionMCDropout(C, None, -1, B)
        model = ResidentReplayModel().to(dev)
        images = [str(i) for i in range(n_total)]
        existing = [[(64 * (i % 3), 128 * (i % 5), 128, 128)] if i % 2 == 0 else [] for i in range(n_total)]
        sel.create_region_maps(model, images[: world * B], existing[: world * B], a.R, a.k)      # warm-up
        bench.barrier(world)
        n0 = _lib.launch_count()
        t0 = time.perf_counter()
        regions, count = sel.create_region_maps(model, images, existing, a.R, a.k)
        torch.cuda.synchronize()
        dt = bench.max_over_ranks(time.perf_counter() - t0, world)
        launches = _lib.launch_count() - n0
    finally:
        base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
    if rank == 0:
        print(json.dumps({
            "workload": f"region_vote_entropy_{H}x{W}_c{C}_t{T}_R{a.R}_k{a.k}" + ("_lowres_logits" if a.lowres else ""),
            "n_gpus": world, "images": n_total, "images_per_rank": a.images_per_rank, "batch": B, "scaling": "weak",
            "seconds": round(dt, 4), "images_per_s": round(n_total / dt, 1), "regions_picked": int(count),
            "images_with_regions": len(regions), "gpu_launches_rank0": int(launches),
            "api": "ActiveSelectionMCDropout.create_region_maps(model, images, existing_regions, 128, 125)",
            "timing": "wall clock around the call, barrier + synchronize on both sides, max over ranks",
        }), flush=True)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()
