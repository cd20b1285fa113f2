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
NUMA_INFO = None
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
    ap.add_argument("--no-sweep", action="store_true", help="skip the pass-group sweep (streaming vs single-shot)")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 3 / 4 / 5 sub-records")
    return ap.parse_args()


def workload_config(args, world):
    """The `config` of the JSON line - identical for both arms (--impl b200 / reference) of the same command line."""
    B, G = args.batch, max(1, min(args.pass_group, T))
    return {"workload": WORKLOAD, "pool_images": POOL_IMAGES, "H": H, "W": W, "classes": C, "mc_passes": T,
            "batch_images_per_step": B, "pass_group": G, "topk": TOPK, "mode": args.mode,
            "scores": {"full": "vote_entropy+pred_entropy+bald+confidence+margin", "votes": "vote_entropy",
                       "probs": "pred_entropy+bald+confidence+margin"}[args.mode],
            "sharding": f"by image, {world} rank(s), candidate all-gather only",
            "l2": f"inputs {T * B * C * H * W * 4 / 1e9:.2f} GB per step >> 126 MB L2 (no flush needed)"}


def peaks_sm_mhz():
    return measured_peaks()[0].get("sm_max_mhz", 1965.0)


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
        if os.environ.get("DAS_NUMA_BIND", "1") != "0":
            # every rank next to its GPU: pinned host buffers are then allocated on the GPU's NUMA node
            from deep_active_semantic_segmentation_b200 import dist as _dist
            global NUMA_INFO
            NUMA_INFO = _dist.bind_to_gpu_numa_node(local)
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

    sweep = None
    if rank == 0 and not args.no_sweep:
        try:
            sweep = [pass_group_sweep(args, dev, peaks, passes, labels, b) for b in sorted({min(2, B), B})]
        except Exception as exc:
            sweep = {"error": repr(exc)[:300]}
    e2e = None if args.no_e2e else run_e2e(args, world, rank, dev)
    configs = None
    if not args.no_configs:
        del passes
        torch.cuda.empty_cache()
        try:
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):        # stdout carries exactly one JSON line
                configs = config_records(args, world, rank, dev, peaks)
        except Exception as exc:  # informative only: never lose the headline line
            import traceback
            traceback.print_exc()
            configs = {"error": repr(exc)[:300]}
        passes = None
    cpu = None
    torch_gpu = None
    fused_up = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if passes is None:
            passes, labels = synth.device_pass_logits(synth.DEFAULT_SEED, rank * K * B, B, T, C, H, W, dev)
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
            "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu, "torch_gpu_baseline": torch_gpu, "e2e": e2e,
            "fused_upsample_variant": fused_up, "pass_group_sweep": sweep, "configs": configs,
            "gpu_launches": int(launches), "clocks": clocks,
            "selected_head": [int(v) for v in chosen[:5].tolist()],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


def _timed_call(fn, world):
    """Wall clock around one API call: barrier + synchronize on both sides, max over ranks (seconds)."""
    import torch
    barrier(world)
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return max_over_ranks(time.perf_counter() - t0, world), out


def config_records(args, world, rank, dev, peaks):
    """Compact records of the other BASELINE configs, through the selector API, at this N (every rank takes part):
    config 3 (region vote entropy + NMS, create_region_maps), config 4 (CEAL scores over MC input-noise passes, Pascal
    shape) and config 5 (core-set k-center, N = 10 000).  Logits are resident in HBM (the network's cost is excluded);
    images and labels come from pinned host memory through the batch feeder; results go back to the host."""
    import torch
    from deep_active_semantic_segmentation_b200 import _lib, constants, synth
    from deep_active_semantic_segmentation_b200.active_selection import (ActiveSelectionCoreSet, ActiveSelectionMCDropout,
                                                                         ActiveSelectionMCNoise, base)
    out = {}
    B = args.batch
    per_rank = int(os.environ.get("DAS_BENCH_SUB_IMAGES", 512))

    def resident_setup(Hs, Ws, Cs, seed):
        passes, labels = synth.device_pass_logits(seed, rank * per_rank, B, T, Cs, Hs, Ws, dev)
        labels_h = torch.empty(labels.shape, dtype=labels.dtype, pin_memory=True).copy_(labels)
        image = torch.zeros(3, Hs, Ws).pin_memory()

        class DS(torch.utils.data.Dataset):
            def __init__(self, env, paths, crop_size, include_labels=False):
                self.paths, self.include_labels = paths, include_labels

            def __len__(self):
                return len(self.paths)

            def __getitem__(self, i):
                return {"image": image, "label": labels_h[int(self.paths[i]) % B]} if self.include_labels else image

        class Model(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.drop = torch.nn.Dropout2d(0.25)
                self.t = 0

            def forward(self, x):
                o = passes[self.t % T]
                self.t += 1
                return o[:x.shape[0]]

        return DS, Model().to(dev)

    old_ds, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
    n_total = world * per_rank
    images = [str(i) for i in range(n_total)]
    try:
        # ---- config 3: region-based vote entropy (128 x 128 regions), pool sharded by image ----
        Hs, Ws, Cs, R_, k_ = 512, 1024, 19, 128, 125
        DS, model = resident_setup(Hs, Ws, Cs, synth.DEFAULT_SEED + 31)
        base.paths_dataset.PathsDataset, constants.MC_STEPS = DS, T
        sel = ActiveSelectionMCDropout(Cs, None, -1, B)
        existing = [[(64 * (i % 3), 128 * (i % 5), 128, 128)] if i % 2 == 0 else [] for i in range(n_total)]
        sel.create_region_maps(model, images[: world * 2 * B], existing[: world * 2 * B], R_, k_)
        n0 = _lib.launch_count()
        dt, (regions, count) = _timed_call(lambda: sel.create_region_maps(model, images, existing, R_, k_), world)
        bytes_img = T * Cs * Hs * Ws * 4
        ach = bytes_img * per_rank / dt / 1e9
        out["config3_region_vote_entropy"] = {
            "workload": f"region_vote_entropy_{Hs}x{Ws}_c{Cs}_t{T}_R{R_}_k{k_}", "value": round(n_total / dt, 1), "unit": UNIT,
            "images": n_total, "seconds": round(dt, 4), "regions_picked": int(count), "gpu_launches_rank0": int(_lib.launch_count() - n0),
            "host_ms_per_batch": round(sel.last_loader.host_seconds / max(sel.last_loader.batches, 1) * 1e3, 4),
            "feeder_path": sel.last_loader.path,
            "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(ach / peaks["hbm_gbs"], 4), "algorithmic_bytes_per_image": bytes_img,
                         "note": "per GPU, whole create_region_maps call (feeder, votes, suppress, box sums, min-max, NMS, merge)"},
            "api": "ActiveSelectionMCDropout.create_region_maps(model, images, existing_regions, 128, 125)"}
        del model
        # ---- config 4: CEAL entropy / margin / confidence over T input-noise passes, Pascal shape ----
        Hs, Ws, Cs, k_ = 513, 513, 21, 60
        DS, model = resident_setup(Hs, Ws, Cs, synth.DEFAULT_SEED + 41)
        base.paths_dataset.PathsDataset, constants.MC_STEPS = DS, T
        seln = ActiveSelectionMCNoise(Cs, None, Hs, B)
        seln.get_mc_scores_for_images_with_input_noise(model, images[: world * 2 * B], k_, score="pred_entropy")
        n0 = _lib.launch_count()
        dt, (chosen, allv) = _timed_call(
            lambda: seln.get_mc_scores_for_images_with_input_noise(model, images, k_, score="pred_entropy"), world)
        bytes_img = T * Cs * Hs * Ws * 4
        ach = bytes_img * per_rank / dt / 1e9
        out["config4_ceal_mc_noise"] = {
            "workload": f"ceal_entropy_margin_mc_input_noise_{Hs}x{Ws}_c{Cs}_t{T}_top{k_}", "value": round(n_total / dt, 1),
            "unit": UNIT, "images": n_total, "seconds": round(dt, 4), "gpu_launches_rank0": int(_lib.launch_count() - n0),
            "host_ms_per_batch": round(seln.last_loader.host_seconds / max(seln.last_loader.batches, 1) * 1e3, 4),
            "scores": "pred_entropy (ranked) + margin + confidence + bald + vote_entropy of the MC-mean softmax",
            "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(ach / peaks["hbm_gbs"], 4), "algorithmic_bytes_per_image": bytes_img,
                         "note": "per GPU, whole selector call; the T noise draws on the image batch are torch ops on the device"},
            "api": "ActiveSelectionMCNoise.get_mc_scores_for_images_with_input_noise(model, images, 60, score='pred_entropy')",
            "selected_head": [int(p) for p in chosen[:5]]}
        del model
        # ---- config 5: core-set k-center greedy, N = 10 000 rows (replicated greedy loop; forwards sharded) ----
        N5, D5, L5, K5 = 10000, 2048, 50, 500
        feats = torch.from_numpy(synth.coreset_features(11, N5, D5)).to(dev)
        cs = ActiveSelectionCoreSet(None, 513, B)
        cs._select_batch(feats, list(range(L5)), 8)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        runs = []
        for _ in range(3):      # the selection is one 5 ms call: best of three, every run listed
            barrier(world)
            ev[0].record()
            picks = cs._select_batch(feats, list(range(L5)), K5)
            ev[1].record()
            torch.cuda.synchronize()
            runs.append(max_over_ranks(ev[0].elapsed_time(ev[1]), world))
        ms = min(runs)
        # the tensor-core part alone: bf16 tcgen05 distance GEMM (2 N^2 Dp flop) incl. the bf16 / norm preparation
        from deep_active_semantic_segmentation_b200 import ops
        ops.KCenterFilter(feats)
        torch.cuda.synchronize()
        gemm_runs = []
        for _ in range(3):      # allocation of the 0.44 GB table + memset + bf16 / norm preparation + the GEMM; best of three
            ev[0].record()
            ops.KCenterFilter(feats)
            ev[1].record()
            torch.cuda.synchronize()
            gemm_runs.append(ev[0].elapsed_time(ev[1]))
        gemm_ms = min(gemm_runs)
        flops = 2.0 * N5 * N5 * (-(-D5 // 64) * 64)
        tf = flops / (gemm_ms * 1e-3) / 1e12
        out["config5_coreset_kcenter"] = {
            "workload": f"coreset_kcenter_N{N5}_D{D5}_L{L5}_K{K5}", "value": round(ms, 3), "unit": "ms per selection",
            "higher_is_better": False, "parallelism": "replicas: every rank runs the whole deterministic greedy loop (DESIGN.md section 6)",
            "runs_ms": [round(v, 3) for v in runs], "filter_build_ms": round(gemm_ms, 4), "greedy_ms": round(ms - gemm_ms, 3),
            "exact_fraction": round(cs.last_filter_stats[0] / max(cs.last_filter_stats[1], 1), 5) if cs.last_filter_stats else None,
            "picks_head": [int(v) for v in picks[:5]],
            "roofline": {"bound": "tensor", "kernel": "kc_dist_gemm2_kernel (tcgen05 cta_group::2, bf16) + kc_prepare_kernel",
                         "achieved": round(tf, 1), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": round(tf / peaks["bf16_tflops"], 4), "flops": flops,
                         "reference_algorithmic_flops": 2.0 * N5 * D5 * (L5 + K5)},
            "api": "ActiveSelectionCoreSet._select_batch(features, selected_indices, 500)"}
        # end to end with the forwards: a stand-in network draws DeepLab-shaped features [B,304,129,129] on the device,
        # the selector avg-pools them to 2736-d rows; each rank forwards ONLY its slice of the images, rows are all-gathered
        Nf = int(os.environ.get("DAS_BENCH_CORESET_IMAGES", 2000))
        image5 = torch.zeros(3, 513, 513).pin_memory()

        class DS5(torch.utils.data.Dataset):
            def __init__(self, env, paths, crop_size, include_labels=False):
                self.paths = paths

            def __len__(self):
                return len(self.paths)

            def __getitem__(self, i):
                return image5

        class FeatModel(torch.nn.Module):
            model_name = "deeplab"

            def __init__(self):
                super().__init__()
                self.gen = torch.Generator(device=dev).manual_seed(5 + 7919 * rank)     # distinct rows on every rank
                self.buf = torch.empty((B, 304, 129, 129), device=dev)

            @property
            def module(self):
                return self

            def set_return_features(self, flag):
                pass

            def forward(self, x):
                f = self.buf[:x.shape[0]].normal_(generator=self.gen)
                return None, f

        base.paths_dataset.PathsDataset = DS5
        fm = FeatModel()
        paths5 = [str(i) for i in range(Nf)]
        cs.get_k_center_greedy_selections(4, fm, paths5[L5:L5 + 4 * world * B], paths5[:L5])
        runs5 = [_timed_call(lambda: cs.get_k_center_greedy_selections(100, fm, paths5[L5:], paths5[:L5]), world)[0]
                 for _ in range(2)]
        dt = min(runs5)
        out["config5_coreset_kcenter"]["with_forwards"] = {
            "images": Nf, "select": 100, "seconds": round(dt, 4), "runs_s": [round(v, 4) for v in runs5],
            "images_forwarded_per_rank": int(cs.last_forward_rows),
            "note": "feature extraction sharded by image over the ranks + all-gather of the pooled 2736-d rows, then the "
                    "replicated greedy loop; stand-in network = one normal_() per batch on the device",
            "api": "ActiveSelectionCoreSet.get_k_center_greedy_selections(100, model, candidates, already_selected)"}
    finally:
        base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
    return out


def pass_group_sweep(args, dev, peaks, passes, labels, B):
    """The streaming form next to the single-shot one: G passes per launch for G in {1, 4, 5, 10, 20} (G = 1 is the
    literal north-star kernel 1: every pass consumed as it arrives, running state in HBM / L2).  Per G: time of a whole
    batch (all its launches), algorithmic fraction of the HBM peak, DRAM bytes per batch from the committed ncu capture."""
    import torch
    from deep_active_semantic_segmentation_b200 import _lib, ops

    sub = [p[:B].contiguous() for p in passes]
    lab = labels[:B]
    traffic = {}
    tf = os.path.join(ROOT, "profiles", "r2_pass_group_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf))
        except Exception:
            traffic = {}
    rows = []
    alg = T * B * C * H * W * 4
    scores = torch.zeros((B, _lib.N_SCORES), dtype=torch.float32, device=dev)
    for G in (1, 4, 5, 10, 20):
        st = ops.MCState(B, C, H, W, T, votes=True, probs=True, device=dev, single_shot=(G >= T))
        groups = [sub[t0:t0 + G] for t0 in range(0, T, G)]

        def batch():
            st.reset()
            for g in groups[:-1]:
                st.accumulate(g)
            st.score(groups[-1], lab, maps=(), scores_out=scores)

        for _ in range(3):
            batch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            batch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ach = alg / (ms * 1e-3) / 1e9
        t_ = traffic.get(f"B{B}_G{G}")
        rows.append({"pass_group": G, "launches_per_batch": len(groups) + 1, "ms_per_batch": round(ms, 4),
                     "images_per_s": round(B / ms * 1e3, 1), "achieved_gbs": round(ach, 1),
                     "frac": round(ach / peaks["hbm_gbs"], 4), "dram_bytes_per_batch": t_,
                     "dram_over_algorithmic": round(t_ / alg, 3) if t_ else None})
    return {"batch_images": B, "algorithmic_bytes_per_batch": alg, "l2_persist": _lib.get_option("mc_l2_persist"),
            "state_bytes": B * (C + 1) * H * W * 4, "l2": _lib.l2_info(dev), "rows": rows}


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

    host_image = torch.zeros(3, H, W).pin_memory()

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
        dt_local = time.perf_counter() - t0
        dt = max_over_ranks(dt_local, world)
    finally:
        base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
    host_ms = sel.last_loader.host_seconds / max(sel.last_loader.batches, 1) * 1e3
    # per-rank host-to-device rate of this run (every rank pulls T*B logits tensors per step through its own PCIe link
    # out of the same host DRAM): names the limiter of the multi-GPU e2e curve
    h2d_step = T * B * C * H * W * 4 + B * H * W * 4 + B * 3 * H * W * 4
    my_gbs = h2d_step * K / dt_local / 1e9
    if world > 1:
        import torch.distributed as td
        rates = [None] * world
        td.all_gather_object(rates, round(my_gbs, 2))
    else:
        rates = [round(my_gbs, 2)]
    # second variant: the same call when the logits of the T stochastic passes are already in HBM (a network on the
    # device produced them; its cost is excluded so that the number shows what the SELECTOR costs): images + labels still
    # come from pinned host memory every step through the batch feeder, scores and the ranking go back to the host
    dev_variant = None
    try:
        Bd, Kd = args.batch, 8 * K
        passes_d, labels_d = synth.device_pass_logits(synth.DEFAULT_SEED + 2, 0, Bd, T, C, H, W, dev)
        host_labels_d = torch.empty(labels_d.shape, dtype=labels_d.dtype, pin_memory=True).copy_(labels_d)
        host_image_p = torch.zeros(3, H, W).pin_memory()

        class PinnedDataset(HostDataset):
            def __getitem__(self, i):
                return {"image": host_image_p, "label": host_labels_d[int(self.paths[i]) % Bd]}

        class ResidentLogitsModel(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.drop = torch.nn.Dropout2d(0.25)
                self.t = 0

            def forward(self, x):
                out = passes_d[self.t % T]
                self.t += 1
                return out[:x.shape[0]]

        old_ds, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
        base.paths_dataset.PathsDataset, constants.MC_STEPS = PinnedDataset, T
        try:
            sel2 = ActiveSelectionMCDropout(C, None, -1, Bd)
            model2 = ResidentLogitsModel().to(dev)
            images2 = [str(i) for i in range(world * Kd * Bd)]
            sel2.get_mc_scores_for_images(model2, images2[: world * 2 * Bd], TOPK)
            barrier(world)
            t0 = time.perf_counter()
            sel2.get_mc_scores_for_images(model2, images2, TOPK)
            torch.cuda.synchronize()
            dt2 = max_over_ranks(time.perf_counter() - t0, world)
            ld = sel2.last_loader
        finally:
            base.paths_dataset.PathsDataset, constants.MC_STEPS = old_ds, old_T
        dev_variant = {"value": round(world * Kd * Bd / dt2, 2), "unit": UNIT, "batch_images_per_step": Bd,
                       "h2d_bytes_per_step": Bd * H * W * 4 + Bd * 3 * H * W * 4, "d2h_bytes_per_step": Bd * 6 * 4 + min(TOPK, Bd) * 16,
                       "steps": Kd, "host_ms_per_batch": round(ld.host_seconds / max(ld.batches, 1) * 1e3, 4),
                       "feeder_path": ld.path,
                       "note": "images + labels from pinned host memory through the batch feeder every step; the T logits "
                               "tensors of a batch are resident in HBM (the network's cost is excluded)"}
        del passes_d
    except Exception as exc:  # the strict variant above is the contract; this one is informative
        dev_variant = {"error": repr(exc)[:200]}
    return {"value": round(world * K * B / dt, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_step, "logits_on_device_variant": dev_variant,
            "d2h_bytes_per_step": B * 6 * 4 + min(TOPK, B) * 16, "steps": K, "batch_images_per_step": B,
            "host_ms_per_batch": round(host_ms, 4), "h2d_gbs_per_rank": rates,
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
    # which kernel the library launches for this shape (csrc/mc_api.cu up_warps(): one pixel per lane for C <= 20 and
    # C >= 22, pixel pairs at C = 21 and for vote-only scoring)
    variant = ops.upsample_variant(B, C, h, w, H, W, votes=votes, probs=probs, device=dev)
    one_pixel = variant >= 100
    out = {"value": round(B / ms * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": steps,
           "kernel": ("mc_score_up1_kernel (one pixel per lane, class pairs in the packed pipe)" if one_pixel else
                      "mc_score_up_kernel (pixel pairs)") + ": fused bilinear upsample + K1 + K2", "variant": variant, "lowres": [h, w],
           "hbm_bytes_per_step": T * B * C * h * w * 4,
           "fullres_bytes_avoided_per_step": 2 * T * B * C * H * W * 4,
           "note": "the network no longer writes T*B*C*H*W*4 bytes of interpolated logits and the scorer no longer reads them"}
    # The kernel reads 16x fewer bytes than the resident-logits kernel (5 % of the HBM peak): its bounds are on chip.
    #   MUFU: per pixel and pass C ex2 + 1 rcp + 1 lg2, per pixel C + 1 lg2 in the finalize; 16 MUFU lanes / clk / SM
    #   issue: ~10 warp instructions per logit (pass loop + producers + finalize, ncu), 4 / clk / SM
    sm_hz = float(peaks_sm_mhz()) * 1e6
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    px = B * H * W
    mufu_ops = px * (T * (C + 2) + (C + 1)) if probs else 0
    mufu_peak = n_sm * 16 * sm_hz
    # warp instructions per launch measured by ncu at B = 8, 512 x 1024, C = 19, T = 20 (profiles/
    # r2_fused_upsample_ncu_summary.txt), scaled linearly in pixels, passes and classes: 703.8 M for the one-pixel-per-lane
    # kernel the library picks for C <= 20 and C >= 22, 601.6 M for the pixel-pair kernel (C = 21; 338.4 M measured on
    # the Pascal shape = the same 0.38 per pixel, pass and class)
    warp_instr = (703.80e6 if one_pixel else 601.57e6) * (px / (8.0 * 512 * 1024)) * (T * C) / (20.0 * 19.0)
    issue_peak = n_sm * 4 * sm_hz
    out["roofline"] = {"bound": "mufu", "algorithmic_ops": int(mufu_ops), "peak_ops_per_s": mufu_peak,
                       "achieved_ops_per_s": round(mufu_ops / (ms * 1e-3), 1), "frac": round(mufu_ops / (ms * 1e-3) / mufu_peak, 4),
                       "mufu_bound_ms": round(mufu_ops / mufu_peak * 1e3, 4),
                       "issue_bound": {"warp_instructions": int(warp_instr), "peak_warp_instr_per_s": issue_peak,
                                       "frac": round(warp_instr / (ms * 1e-3) / issue_peak, 4),
                                       "issue_bound_ms": round(warp_instr / issue_peak * 1e3, 4)},
                       "note": "MUFU = 16 lanes/clk/SM at the maximum SM clock; the issue bound (one warp instruction per "
                               "scheduler and clock) is the tighter one for this kernel: see profiles/r2_upsample_notes.md"}
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
            "config": workload_config(args, max(1, int(args.gpus))),     # the same dict as the b200 arm of this N
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
