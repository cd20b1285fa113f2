#!/usr/bin/env python
"""Benchmark of the active-selection scoring hot path (BASELINE.json metric):
pool images scored / s for T Monte-Carlo passes (vote entropy + predictive entropy + BALD + top-k).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE config 2: Cityscapes-shaped pool, 512x1024, C=19, T=20.  One step = one batch of
B pool images through the whole path: T passes of logits -> K1 accumulate (pass groups of G) -> K2
finalize (image scores written straight into the pool score table); after the K steps one K3 top-k
(+ the candidate all-gather for N > 1) - all inside the timed region.  Inputs are synthetic logits that
are resident in HBM before the timed region starts (T*B*C*H*W*4 bytes per step, far larger than L2).

`--impl reference` times the reference's own CPU op sequence (oracle/cpu_port.py, all host threads) on a
bounded sample of the same workload; the reference is pure Python and cannot be installed on the box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# NCCL prints its version banner on stdout when NCCL_DEBUG is VERSION: keep stdout to the one JSON line
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, C, T = 512, 1024, 19, 20
POOL_IMAGES, TOPK = 2975, 125
METRIC = "pool images scored/sec (T MC passes, entropy+BALD+top-k)"
UNIT = "images/s"
WORKLOAD = "mc_dropout_entropy_bald_cityscapes_pool_512x1024_c19_t20"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("DAS_BENCH_BATCH", 8)))
    ap.add_argument("--pass-group", type=int, default=int(os.environ.get("DAS_BENCH_PASS_GROUP", 20)))
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-batch", type=int, default=2)
    ap.add_argument("--mode", default="full", choices=["full", "votes", "probs"],
                    help="full = vote entropy + softmax scores (default); votes = the reference's vote entropy only; "
                         "probs = softmax-mean entropy / BALD / confidence / margin only")
    ap.add_argument("--workload", default="cityscapes", choices=["cityscapes", "pascal"],
                    help="cityscapes = BASELINE config 2 (default, the metric's config); pascal = config 4 shape: "
                         "513x513, C=21, T=20, pool 10582, top-60 (odd planes: flat 1-D TMA maps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-upsample-variant", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin: float, t_end: float):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [l for (t, l) in self.lines if t_begin <= t <= t_end + 0.2] or [l for _, l in self.lines]
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                smax = max(smax, float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def init_dist(n_gpus: int):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as td
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the first communicator is created; stdout must carry
        # exactly one JSON line, so the banner is sent to stderr (fd level: it comes from C code)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            td.init_process_group("nccl", device_id=torch.device("cuda", local))
            td.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as td
        td.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    import torch
    import torch.distributed as td
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    from deep_active_semantic_segmentation_b200 import _lib, dist, ops, synth
    from deep_active_semantic_segmentation_b200._lib import SCORE_INDEX

    world, rank, local = init_dist(args.gpus)
    dev = torch.device("cuda", local)
    B, G, K, Wm = args.batch, max(1, min(args.pass_group, T)), args.steps, max(args.warmup, 3)
    lib = _lib.load(build_if_missing=False)      # fail loudly if the CUDA library is missing

    # synthetic pool slice, resident in HBM: T pass buffers of [B,C,H,W] (+ labels); every step re-reads
    # T*B*C*H*W*4 bytes (6.4 GB at B=8) >> 126 MB L2, so the logits always come from HBM
    passes, labels = synth.device_pass_logits(synth.DEFAULT_SEED, rank * K * B, B, T, C, H, W, dev)
    votes, probs = args.mode in ("full", "votes"), args.mode in ("full", "probs")
    state = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, device=dev, single_shot=(G >= T))
    pool_scores = torch.zeros((K * B, _lib.N_SCORES), dtype=torch.float32, device=dev)
    groups = [passes[t0:t0 + G] for t0 in range(0, T, G)]
    acc_events, fin_events = [], []

    def step(i, record):
        # all but the last pass group: K1 (streaming accumulate); last group: fused K1+K2, which writes the
        # image scores of this batch straight into the pool score table
        state.reset()
        for gi, grp in enumerate(groups):
            last = gi == len(groups) - 1
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if last:
                state.score(grp, labels, maps=(), scores_out=pool_scores[i * B:(i + 1) * B])
            else:
                state.accumulate(grp)
            if record:
                e1.record()
                (fin_events if last else acc_events).append((e0, e1, len(grp)))

    def select():
        # K3 on the shard -> candidate records -> ONE all-gather -> K3 merge on every rank -> one D2H of the winners
        col = pool_scores[:, SCORE_INDEX["bald" if probs else "vote_entropy"]].contiguous()
        return dist.select_ranked(col, min(TOPK, world * col.numel()), True, id_offset=rank * K * B)[1]

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(Wm):
        step(i % K, False)
    select()
    barrier(world)
    launches0 = _lib.launch_count()
    t_begin = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()      # `ncu --profile-from-start off` then lists the timed region only
    start.record()
    for i in range(K):
        step(i, True)
    chosen = select()
    end.record()
    torch.cuda.profiler.stop()
    barrier(world)
    t_end = time.perf_counter()
    launches = _lib.launch_count() - launches0
    elapsed_ms = max_over_ranks(start.elapsed_time(end), world)
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    value = world * B * K / (elapsed_ms * 1e-3)

    # dominant kernel: K1 mc_accumulate.  algorithmic bytes per launch = G passes * B*C*H*W*4 (each logit once)
    # dominant kernel = the MC kernel (K1 / fused K1+K2); algorithmic bytes of a launch = its passes * B*C*H*W*4
    ev = fin_events + acc_events
    k_ms = [a.elapsed_time(b) for a, b, _ in ev]
    k_bytes = [n * B * C * H * W * 4 for _, _, n in ev]
    avg_acc_ms = sum(k_ms) / len(k_ms)
    alg_bytes = sum(k_bytes) / len(k_bytes)
    peaks, peak_kind = measured_peaks()
    achieved = sum(k_bytes) / (sum(k_ms) * 1e-3) / 1e9
    tma = G >= T and _lib.get_option("mc_tma") == 1
    traffic = None
    tf = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tf):
        try:
            tj = json.load(open(tf))
            if tj.get("batch") == B and tj.get("pass_group") == G and tj.get("tma", False) == tma and tj.get("mode", "full") == args.mode:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    kname = ("mc_score_tma_kernel<C=19> (fused K1+K2, TMA ring, persistent)" if tma else
             "mc_score_kernel<C=19,VEC=2> (fused K1+K2, LDG)" if G >= T else
             "mc_accumulate_kernel / mc_score_kernel <C=19,VEC=2>")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1),
                "peak": peaks["hbm_gbs"], "peak_kind": peak_kind + " copy bandwidth (burst)", "unit": "GB/s",
                "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": int(alg_bytes), "avg_launch_ms": round(avg_acc_ms, 4),
                "launches_timed": len(k_ms), "kernel_share_of_step": round(sum(k_ms) / elapsed_ms, 4)}

    e2e = None if args.no_e2e else run_e2e(args, world, rank, dev)
    cpu = None
    torch_gpu = None
    fused_up = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline([p[:1].cpu() for p in passes], labels[:1].cpu(), budget_s=20.0)
        torch_gpu = torch_gpu_baseline(passes, labels)
    if rank == 0 and world == 1 and not args.no_upsample_variant:
        try:
            fused_up = fused_upsample_variant(args, dev)
        except Exception as exc:  # informative only: never lose the headline line
            fused_up = {"error": repr(exc)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": round(elapsed_ms / K, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pool_images": POOL_IMAGES, "H": H, "W": W, "classes": C, "mc_passes": T,
                       "batch_images_per_step": B, "pass_group": G, "topk": TOPK, "mode": args.mode,
                       "scores": {"full": "vote_entropy+pred_entropy+bald+confidence+margin", "votes": "vote_entropy",
                                  "probs": "pred_entropy+bald+confidence+margin"}[args.mode],
                       "sharding": f"by image, {world} rank(s), candidate all-gather only",
                       "l2": f"inputs {T * B * C * H * W * 4 / 1e9:.2f} GB per step >> 126 MB L2 (no flush needed)"},
            "roofline": roofline, "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "e2e": e2e,
            "fused_upsample_variant": fused_up, "gpu_launches": int(launches), "clocks": clocks,
            "selected_head": [int(v) for v in chosen[:5].tolist()],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


def run_e2e(args, world, rank, dev):
    """Same metric through the public selector API with HOST logits: every step copies T*B logits
    tensors from pinned host memory to the device and the ranking result back to the host."""
    import torch
    from deep_active_semantic_segmentation_b200 import constants, synth
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCDropout, base

    B, K = args.e2e_batch, args.e2e_steps
    passes, labels = synth.device_pass_logits(synth.DEFAULT_SEED + 1, 0, B, T, C, H, W, dev)
    host = [torch.empty(p.shape, dtype=p.dtype, pin_memory=True).copy_(p) for p in passes]
    host_labels = torch.empty(labels.shape, dtype=labels.dtype, pin_memory=True).copy_(labels)
    del passes
    torch.cuda.synchronize()

    host_image = torch.zeros(3, H, W)

    class HostDataset(torch.utils.data.Dataset):   # stands in for PathsDataset: image + label from host memory
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            return {"image": host_image, "label": host_labels[int(self.paths[i]) % B]}

    class HostReplayModel(torch.nn.Module):      # the network forward is out of scope: logits arrive from the host
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.t = 0

        def forward(self, x):
            out = host[self.t % T].to(dev, non_blocking=True)
            self.t += 1
            return out[:x.shape[0]]

    old_ds, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
    base.paths_dataset.PathsDataset, constants.MC_STEPS = HostDataset, T
    try:
        sel = ActiveSelectionMCDropout(C, None, -1, B)
        # default pass grouping: the T copies of a batch are enqueued back to back, then ONE fused launch scores them
        model = HostReplayModel().to(dev)
        images = [str(i) for i in range(world * K * B)]
        sel.get_mc_scores_for_images(model, images[: world * B], TOPK)      # warm-up
        barrier(world)
        t0 = time.perf_counter()
        chosen, _ = sel.get_mc_scores_for_images(model, images, TOPK)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, world)
    finally:
        base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
    # second variant: the same call with a stand-in network that lives on the device (images + labels still come
    # from pinned host memory every step; the logits of each stochastic pass are produced on the GPU, as a real
    # forward would) - what the selector costs when PCIe does not carry the logits
    dev_variant = None
    try:
        base_logits, _ = synth.device_pass_logits(synth.DEFAULT_SEED + 2, 0, B, 1, C, H, W, dev)
        base_logits = base_logits[0]
        bufs = [torch.empty_like(base_logits) for _ in range(T)]

        class DeviceNoiseModel(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.drop = torch.nn.Dropout2d(0.25)
                self.t = 0

            def forward(self, x):
                out = bufs[self.t % T]
                self.t += 1
                out.normal_(0.0, 0.7).add_(base_logits)      # dropout-like jitter around the deterministic logits
                return out[:x.shape[0]]

        host_image_p = torch.zeros(3, H, W).pin_memory()

        class PinnedDataset(HostDataset):
            def __getitem__(self, i):
                return {"image": host_image_p, "label": host_labels[int(self.paths[i]) % B]}

        old_ds, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
        base.paths_dataset.PathsDataset, constants.MC_STEPS = PinnedDataset, T
        try:
            sel2 = ActiveSelectionMCDropout(C, None, -1, B)
            model2 = DeviceNoiseModel().to(dev)
            K2 = 4 * K
            images2 = [str(i) for i in range(world * K2 * B)]
            sel2.get_mc_scores_for_images(model2, images2[: world * B], TOPK)
            barrier(world)
            t0 = time.perf_counter()
            sel2.get_mc_scores_for_images(model2, images2, TOPK)
            torch.cuda.synchronize()
            dt2 = max_over_ranks(time.perf_counter() - t0, world)
        finally:
            base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
        dev_variant = {"value": round(world * K2 * B / dt2, 2), "unit": UNIT,
                       "h2d_bytes_per_step": B * H * W * 4 + B * 3 * H * W * 4, "d2h_bytes_per_step": B * 6 * 4 + min(TOPK, B) * 12,
                       "steps": K2, "note": "images + labels from pinned host memory; a stand-in network on the device draws "
                                            "the T stochastic logits (torch normal_ + add_, ~3x the scoring kernel's HBM traffic)"}
        del bufs, base_logits
    except Exception as exc:  # the strict variant above is the contract; this one is informative
        dev_variant = {"error": repr(exc)[:200]}
    h2d = T * B * C * H * W * 4 + B * H * W * 4 + B * 3 * H * W * 4
    return {"value": round(world * K * B / dt, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "logits_on_device_variant": dev_variant,
            "d2h_bytes_per_step": B * 6 * 4 + min(TOPK, B) * 12, "steps": K, "batch_images_per_step": B,
            "api": "ActiveSelectionMCDropout.get_mc_scores_for_images(model, images, k)",
            "note": "logits for every pass copied from pinned host memory (PCIe bound)"}


def fused_upsample_variant(args, dev, steps=60):
    """SURVEY 8(f)-1, reported next to the headline (rank 0, N = 1): the same pool step when the network hands over
    its LOW-RESOLUTION decoder logits [B,C,H/4,W/4] (models/deeplab.py:58) and the final bilinear upsample
    (models/deeplab.py:59) runs inside the scoring kernel - device-resident, and end to end through the selector
    with the low-resolution logits of every pass coming from pinned host memory."""
    import torch
    from deep_active_semantic_segmentation_b200 import _lib, constants, ops, synth
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCDropout, base

    B = args.batch
    h, w = (H + 3) // 4, (W + 3) // 4               # DeepLab's stride-4 decoder: 128 x 256 / 129 x 129
    if not ops.upsample_supported(h, w, H, W):
        return {"unavailable": f"{h}x{w} -> {H}x{W} is outside the fused kernel's range"}
    low, lab_low = synth.device_pass_logits(synth.DEFAULT_SEED + 3, 0, B, T, C, h, w, dev, block=8)
    labels = torch.nn.functional.interpolate(lab_low[:, None], size=(H, W), mode="nearest")[:, 0].contiguous()
    votes, probs = args.mode in ("full", "votes"), args.mode in ("full", "probs")
    st = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, device=dev, single_shot=True)
    scores = torch.zeros((B, _lib.N_SCORES), dtype=torch.float32, device=dev)

    def step():
        st.reset()
        st.score_upsampled(low, labels, maps=(), scores_out=scores)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": round(B / ms * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": steps,
           "kernel": "mc_score_up_kernel (fused bilinear upsample + K1 + K2)", "lowres": [h, w],
           "bound": "issue slots / latency (not HBM)", "hbm_bytes_per_step": T * B * C * h * w * 4,
           "fullres_bytes_avoided_per_step": 2 * T * B * C * H * W * 4,
           "note": "the network no longer writes T*B*C*H*W*4 bytes of interpolated logits and the scorer no longer reads them"}
    if args.no_e2e:
        return out
    # end to end: low-resolution logits of every pass from pinned host memory through the selector API
    Be, Ke = B, 4 * args.e2e_steps          # 8 images per step: 0.47 GB of low-resolution logits over PCIe
    host = [torch.empty((Be,) + tuple(p.shape[1:]), dtype=p.dtype, pin_memory=True).copy_(p[:Be]) for p in low]
    host_labels = torch.empty((Be, H, W), dtype=torch.float32, pin_memory=True).copy_(labels[:Be])
    host_image = torch.zeros(3, H, W).pin_memory()

    class HostDataset(torch.utils.data.Dataset):
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            return {"image": host_image, "label": host_labels[int(self.paths[i]) % Be]}

    class HostLowResModel(torch.nn.Module):       # a forward that stops at `low_res_x`; its values arrive from the host
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.t = 0

        def forward(self, x):
            out = host[self.t % T].to(dev, non_blocking=True)
            self.t += 1
            return out[:x.shape[0]]

    old_ds, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
    base.paths_dataset.PathsDataset, constants.MC_STEPS = HostDataset, T
    try:
        sel = ActiveSelectionMCDropout(C, None, -1, Be)
        model = HostLowResModel().to(dev)
        images = [str(i) for i in range(Ke * Be)]
        sel.get_mc_scores_for_images(model, images[:Be], TOPK)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sel.get_mc_scores_for_images(model, images, TOPK)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
    out["e2e"] = {"value": round(Ke * Be / dt, 2), "unit": UNIT, "steps": Ke, "batch_images_per_step": Be,
                  "h2d_bytes_per_step": T * Be * C * h * w * 4 + Be * H * W * 4 + Be * 3 * H * W * 4,
                  "d2h_bytes_per_step": Be * 6 * 4 + min(TOPK, Be) * 12,
                  "api": "ActiveSelectionMCDropout.get_mc_scores_for_images(model, images, k), model returns low_res_x",
                  "note": "low-resolution logits of every pass, images and labels copied from pinned host memory"}
    return out


def cpu_baseline(pass_logits_1img, labels_1img, budget_s: float):
    """oracle/cpu_port.py timed on the host cores on a bounded sample: whole images of the same workload."""
    import torch
    from oracle import cpu_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    cpu_port.score_batch(pass_logits_1img, labels_1img, C)
    first = time.perf_counter() - t0
    n = max(1, min(8, int(budget_s / max(first, 1e-3)) - 1))
    t0 = time.perf_counter()
    for _ in range(n):
        cpu_port.score_batch(pass_logits_1img, labels_1img, C)
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} image(s) of {H}x{W}, C={C}, T={T} (after 1 untimed), oracle/cpu_port.py (torch CPU ops in the reference's order)"}


def torch_gpu_baseline(passes, labels):
    """The reference's op sequence (oracle/cpu_port.py, same ATen ops in the same order) run by PyTorch eager on the
    SAME B200, inputs resident in HBM: the "same-box PyTorch" comparator of SURVEY.md section 8(d).  Reported only."""
    import torch
    from oracle import cpu_port

    B = passes[0].shape[0]
    cpu_port.score_batch(passes, labels, C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        cpu_port.score_batch(passes, labels, C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": round(B / (ms * 1e-3), 2), "unit": UNIT, "ms_per_batch": round(ms, 3), "batch_images": B,
            "kind": "reference op sequence in PyTorch eager on this GPU (oracle/cpu_port.py with CUDA tensors)"}


def run_reference(args):
    """Reference arm: the reference's CPU scoring op sequence (port) on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from deep_active_semantic_segmentation_b200 import synth
    from oracle import cpu_port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, Wm = args.steps, max(args.warmup, 1)
    g = torch.Generator().manual_seed(synth.DEFAULT_SEED)
    # bounded sample: a horizontal strip of one pool image per step, sized so the run ends in ~2 minutes
    def make(rows):
        base = torch.randn((1, C, rows, W), generator=g)
        cm = torch.randint(0, C, (1, 1, -(-rows // 32), W // 32), generator=g).repeat_interleave(32, 2).repeat_interleave(32, 3)[:, :, :rows]
        base.scatter_add_(1, cm, torch.full((1, 1, rows, W), 3.0))
        passes = [base + 0.7 * torch.randn((1, C, rows, W), generator=g) for _ in range(T)]
        return passes, cm[0].to(torch.float32)
    passes, lab = make(32)
    t0 = time.perf_counter()
    cpu_port.score_batch(passes, lab, C)
    per_row = (time.perf_counter() - t0) / 32
    rows = H
    budget_s = float(os.environ.get("DAS_REF_BUDGET_S", 120.0))     # whole --impl reference run, seconds
    while rows > 16 and per_row * rows * (K + Wm) > budget_s:
        rows //= 2
    passes, lab = make(rows)
    for _ in range(Wm):
        cpu_port.score_batch(passes, lab, C)
    t0 = time.perf_counter()
    for _ in range(K):
        cpu_port.score_batch(passes, lab, C)
    dt = time.perf_counter() - t0
    frac = rows / H
    value = K * frac / dt
    sample = f"{K} steps x ({rows}/{H} rows of one {H}x{W} image, C={C}, T={T}); oracle/cpu_port.py, torch CPU ops in the reference's order"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": Wm, "ms_per_step": round(dt / K * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pool_images": POOL_IMAGES, "H": H, "W": W, "classes": C, "mc_passes": T},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.workload == "pascal":
        H, W, C, T = 513, 513, 21, 20
        POOL_IMAGES, TOPK = 10582, 60
        WORKLOAD = "ceal_mc_noise_pascal_pool_513x513_c21_t20"
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
