#!/usr/bin/env python
"""Where does the host time of the selector-API path go?  Runs get_mc_scores_for_images with a stand-in network that
produces its logits on the device (images + labels still come from pinned host memory through the batch feeder) and
prints images/s, the per-batch host time of the feeder, and a cProfile of the calling thread.

    python tools/probe_selector.py [--batch 8] [--batches 32] [--profile]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--batches", type=int, default=32)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--mode", default="mc", choices=["mc", "region"])
    ap.add_argument("--model", default="resident", choices=["resident", "noise"])
    a = ap.parse_args()
    from deep_active_semantic_segmentation_b200 import constants, synth
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCDropout, base

    H, W, C, T, B = 512, 1024, 19, 20, a.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    passes, labels = synth.device_pass_logits(1, 0, B, T, C, H, W, dev)
    host_labels = torch.empty(labels.shape, dtype=labels.dtype, pin_memory=True).copy_(labels)
    host_image = torch.zeros(3, H, W).pin_memory()

    class DS(torch.utils.data.Dataset):
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            return {"image": host_image, "label": host_labels[int(self.paths[i]) % B]}

    class Resident(torch.nn.Module):        # logits already in HBM: isolates the selector's own cost
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.t = 0

        def forward(self, x):
            out = passes[self.t % T]
            self.t += 1
            return out[:x.shape[0]]

    class Noise(torch.nn.Module):           # draws the logits of every pass on the device (bench.py's device variant)
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.t = 0
            self.stamps = []

        def forward(self, x):
            if self.t % T == 0:
                self.stamps.append(time.perf_counter())
            out = passes[self.t % T]
            self.t += 1
            out.normal_(0.0, 0.7).add_(passes[(self.t + 3) % T])
            return out[:x.shape[0]]

    base.paths_dataset.PathsDataset, constants.MC_STEPS = DS, T
    sel = ActiveSelectionMCDropout(C, None, -1, B)
    model = (Resident() if a.model == "resident" else Noise()).to(dev)
    images = [str(i) for i in range(a.batches * B)]

    def call():
        if a.mode == "mc":
            return sel.get_mc_scores_for_images(model, images, 125)
        return sel.create_region_maps(model, images, [[] for _ in images], 128, 125)

    sel.get_mc_scores_for_images(model, images[:2 * B], 125)
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"[{a.mode}] rep {rep}: {len(images) / dt:8.1f} images/s  ({dt / a.batches * 1e3:.3f} ms per batch of {B})", flush=True)
        if a.model == "noise":
            st = model.stamps[-a.batches:]
            print("   host ms between batches:", [round((b - a_) * 1e3, 2) for a_, b in zip(st, st[1:])][:24], flush=True)
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        call()
        torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
        print(s.getvalue())


if __name__ == "__main__":
    main()
