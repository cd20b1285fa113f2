"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the reference's own selector
classes (unmodified, on CPU, through oracle/ref_shim.py) on seeded synthetic inputs.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden            # all fixtures
    python -m oracle.gen_golden mc_small   # one fixture

Inputs are regenerated at test time from (seed, shape) by
deep_active_semantic_segmentation_b200/synth.py, so only the reference's OUTPUTS (and an
input checksum) are stored; the two NMS PNG fixtures of the reference's own test
(active_selection/tests.py:213-231) are stored as uint8 arrays.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from deep_active_semantic_segmentation_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def checksum(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def versions():
    import sklearn
    import torch

    return np.array([f"torch={torch.__version__}", f"numpy={np.__version__}", f"sklearn={sklearn.__version__}"])


class SortedCapture:
    """Replaces the builtin `sorted` inside a reference module so that the (score, path)
    pairs the selector ranks can be recorded (the selectors only return the chosen paths)."""

    def __init__(self):
        self.calls = []

    def __call__(self, iterable, key=None, reverse=False):
        items = list(iterable)
        self.calls.append(([float(x[0]) for x in items], reverse))
        return sorted(items, key=key, reverse=reverse)


def paths_to_idx(paths):
    return np.array([int(p) for p in paths], dtype=np.int64)


def regions_to_array(regions: dict, N: int):
    rows = []
    for i in range(N):
        for (r, c, h, w) in regions.get(str(i), []):
            rows.append((i, r, c, h, w))
    return np.array(rows, dtype=np.int64).reshape(-1, 5)


def make_pool(seed, N, T_total, C, H, W, block):
    gs = list(range(N))
    logits = synth.pool_logits(seed, gs, T_total, C, H, W, block)
    labels = synth.pool_labels(seed, gs, H, W, C, block)
    return ref_shim.SyntheticPool(logits, labels)


# --------------------------------------------------------------------------------------

def gen_mc(name, seed, N, T, C, H, W, block, k, batch_size):
    """MC-dropout vote entropy (maps + image ranking) and the three CEAL scorers."""
    ref = ref_shim.load_reference()
    torch = ref.torch
    ref.constants.MC_STEPS = T
    pool = make_pool(seed, N, T, C, H, W, block)
    paths = [str(i) for i in range(N)]
    crop = H if H == W else -1

    sel = ref.active_selection.get_active_selection_class("variance", C, pool, crop, batch_size)
    cap = SortedCapture()
    ref.mc_dropout.sorted = cap
    chosen = sel.get_vote_entropy_for_images(ref_shim.make_replay_model(pool), paths, k)
    del ref.mc_dropout.sorted
    ve_scores = np.array(cap.calls[0][0], dtype=np.float32)

    # per-pixel maps of the first batch, straight from _get_vote_entropy_for_batch
    nb = min(batch_size, N)
    ds = ref_shim.SyntheticPathsDataset(pool, paths[:nb], crop, include_labels=True)
    image_batch = torch.stack([ds[i]["image"] for i in range(nb)])
    label_batch = torch.stack([ds[i]["label"] for i in range(nb)])
    maps = sel._get_vote_entropy_for_batch(ref_shim.make_replay_model(pool), image_batch, label_batch)
    ve_maps = np.stack([m.numpy() for m in maps]).astype(np.float32)

    ceal = ref.active_selection.get_active_selection_class("ceal_entropy", C, pool, crop, batch_size)
    ent_sel, ent = ceal.get_maximum_entropy_samples(ref_shim.make_replay_model(pool), paths, k)
    cap = SortedCapture()
    ref.ceal.sorted = cap
    conf_sel = ceal.get_least_confident_samples(ref_shim.make_replay_model(pool), paths, k)
    marg_sel = ceal.get_least_margin_samples(ref_shim.make_replay_model(pool), paths, k)
    del ref.ceal.sorted
    conf_scores = np.array(cap.calls[0][0], dtype=np.float32)
    marg_scores = np.array(cap.calls[1][0], dtype=np.float32)
    weak = ceal.get_weakly_labeled_data(ref_shim.make_replay_model(pool), paths, float(np.median(ent)), entropies=list(ent))

    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, T, C, H, W, block, k, batch_size], dtype=np.int64),
        versions=versions(),
        logits_sha=np.array(checksum(pool.logits)), labels_sha=np.array(checksum(pool.labels)),
        ve_scores=ve_scores, ve_selected=paths_to_idx(chosen), ve_maps=ve_maps,
        ceal_entropy=np.array(ent, dtype=np.float32), ceal_entropy_selected=paths_to_idx(ent_sel),
        ceal_conf=conf_scores, ceal_conf_selected=paths_to_idx(conf_sel),
        ceal_margin=marg_scores, ceal_margin_selected=paths_to_idx(marg_sel),
        weak_idx=paths_to_idx(list(weak.keys())),
        weak_labels=np.stack([weak[p] for p in weak]).astype(np.uint8) if weak else np.zeros((0, H, W), np.uint8),
        weak_threshold=np.float64(np.median(ent)),
    )
    print(f"[golden] {name}: ve_scores[:4]={ve_scores[:4]} selected={paths_to_idx(chosen)}")


def gen_upsample(name, seed, N, T, C, h, w, H, W, block, k, batch_size):
    """The reference selectors on a model that ends like models/deeplab.py:58-59: low-resolution decoder logits
    -> F.interpolate(bilinear, align_corners=True) -> MC vote entropy / CEAL scorers.  Stores the reference's
    outputs and three interpolated class planes (0, C//2, C-1) of image 0, pass 0 (pins oracle.restate.bilinear_upsample_align_corners)."""
    ref = ref_shim.load_reference()
    torch = ref.torch
    ref.constants.MC_STEPS = T
    gs = list(range(N))
    low = synth.pool_logits(seed, gs, T, C, h, w, block)
    labels = synth.pool_labels(seed, gs, H, W, C, block * 4)
    pool = ref_shim.LowResPool(low, labels, H, W)
    paths = [str(i) for i in range(N)]
    crop = H if H == W else -1

    sel = ref.active_selection.get_active_selection_class("variance", C, pool, crop, batch_size)
    cap = SortedCapture()
    ref.mc_dropout.sorted = cap
    chosen = sel.get_vote_entropy_for_images(ref_shim.make_replay_model(pool), paths, k)
    del ref.mc_dropout.sorted
    ve_scores = np.array(cap.calls[0][0], dtype=np.float32)
    nb = min(batch_size, N)
    ds = ref_shim.SyntheticPathsDataset(pool, paths[:nb], crop, include_labels=True)
    image_batch = torch.stack([ds[i]["image"] for i in range(nb)])
    label_batch = torch.stack([ds[i]["label"] for i in range(nb)])
    maps = sel._get_vote_entropy_for_batch(ref_shim.make_replay_model(pool), image_batch, label_batch)
    ve_maps = np.stack([m.numpy() for m in maps]).astype(np.float32)

    ceal = ref.active_selection.get_active_selection_class("ceal_entropy", C, pool, crop, batch_size)
    ent_sel, ent = ceal.get_maximum_entropy_samples(ref_shim.make_replay_model(pool), paths, k)
    cap = SortedCapture()
    ref.ceal.sorted = cap
    conf_sel = ceal.get_least_confident_samples(ref_shim.make_replay_model(pool), paths, k)
    marg_sel = ceal.get_least_margin_samples(ref_shim.make_replay_model(pool), paths, k)
    del ref.ceal.sorted
    up00 = torch.nn.functional.interpolate(torch.from_numpy(low[0, 0][None]), size=(H, W), mode="bilinear",
                                           align_corners=True)[0].numpy()
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, T, C, h, w, H, W, block, k, batch_size], dtype=np.int64),
        versions=versions(), lowres_sha=np.array(checksum(low)), labels_sha=np.array(checksum(labels)),
        upsampled_image0_pass0_classes=up00[[0, C // 2, C - 1]].astype(np.float32),  # 3 class planes keep the file small
        ve_scores=ve_scores, ve_selected=paths_to_idx(chosen), ve_maps=ve_maps,
        ceal_entropy=np.array(ent, dtype=np.float32), ceal_entropy_selected=paths_to_idx(ent_sel),
        ceal_conf=np.array(cap.calls[0][0], dtype=np.float32), ceal_conf_selected=paths_to_idx(conf_sel),
        ceal_margin=np.array(cap.calls[1][0], dtype=np.float32), ceal_margin_selected=paths_to_idx(marg_sel),
    )
    print(f"[golden] {name}: ve_scores[:4]={ve_scores[:4]} selected={paths_to_idx(chosen)}")


def gen_region(name, seed, N, T, C, S, block, R, selection_size, batch_size):
    """create_region_maps of mc_dropout (square crop S), with labelled-region suppression."""
    ref = ref_shim.load_reference()
    ref.constants.MC_STEPS = T
    pool = make_pool(seed, N, T, C, S, S, block)
    paths = [str(i) for i in range(N)]
    rng = np.random.default_rng(seed + 1)
    existing = []
    for i in range(N):
        if i % 3 == 0:
            existing.append([])
        elif i % 3 == 1:
            r0, c0 = (int(v) for v in rng.integers(0, S - R, size=2))
            existing.append([(r0, c0, R, R)])
        else:
            existing.append([(int(rng.integers(0, S - R)), int(rng.integers(0, S - R)), R, R) for _ in range(2)])
    if N > 4:
        existing[4] = [(0, 0, S, S)]  # fully labelled image (region_cityscapes.py:26)

    recorded = {}
    orig = ref.mc_dropout.ActiveSelectionMCDropout.square_nms

    def recording_nms(score_maps, region_size, max_selection_count):
        recorded["norm"] = score_maps.numpy().copy()
        recorded["K"] = float(max_selection_count)
        return orig(score_maps, region_size, max_selection_count)

    ref.mc_dropout.ActiveSelectionMCDropout.square_nms = staticmethod(recording_nms)
    try:
        sel = ref.active_selection.get_active_selection_class("variance", C, pool, S, batch_size)
        regions, count = sel.create_region_maps(ref_shim.make_replay_model(pool), paths, existing, R, selection_size)
    finally:
        ref.mc_dropout.ActiveSelectionMCDropout.square_nms = staticmethod(orig)

    ex_rows = np.array([(i, *rc) for i, lst in enumerate(existing) for rc in lst], dtype=np.int64).reshape(-1, 5)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, T, C, S, block, R, selection_size, batch_size], dtype=np.int64),
        versions=versions(), logits_sha=np.array(checksum(pool.logits)),
        existing=ex_rows, regions=regions_to_array(regions, N), count=np.int64(count),
        norm_maps=recorded["norm"].astype(np.float32), K=np.float64(recorded["K"]),
    )
    print(f"[golden] {name}: {count} regions over {len(regions)} images, K={recorded['K']:.2f}")


def unet_head(seed, N, S):
    """Synthetic 2-channel error-predictor logits [N,2,S,S]: smooth blobs + noise (deterministic)."""
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 911]))
    coarse = rng.standard_normal(size=(N, 2, -(-S // 8), -(-S // 8)), dtype=np.float32) * np.float32(2.0)
    up = np.repeat(np.repeat(coarse, 8, axis=2), 8, axis=3)[:, :, :S, :S]
    return (up + rng.standard_normal(size=(N, 2, S, S), dtype=np.float32)).astype(np.float32)


def gen_accuracy(name, seed, N, C, S, block, R, k, batch_size):
    """ActiveSelectionAccuracy (accuracy.py): label-based error count, softmax / argmax error mass of the
    2-channel error head, 'unsure' score and the least-accurate region maps."""
    ref = ref_shim.load_reference()
    torch = ref.torch
    from active_selection import accuracy as ref_accuracy  # reference module

    pool = make_pool(seed, N, 1, C, S, S, block)
    unet = unet_head(seed, N, S)
    paths = [str(i) for i in range(N)]

    class PairModel(torch.nn.Module):          # returns the segmentation logits or (deeplab_output, unet_output)
        def __init__(self, pair):
            super().__init__()
            self.pair = pair

        @property
        def module(self):
            return self

        def forward(self, x):
            gs = [int(round(float(v) / ref_shim.GID_SCALE)) for v in x[:, 0, 0, 0]]
            seg = torch.from_numpy(np.stack([pool.logits[g, 0] for g in gs]))
            if not self.pair:
                return seg
            return seg, torch.from_numpy(np.stack([unet[g] for g in gs]))

    sel = ref.active_selection.get_active_selection_class("accuracy_labels", C, pool, S, batch_size)
    out = {}
    for key, call in (("labels", lambda: sel.get_least_accurate_sample_using_labels(PairModel(False), paths, k)),
                      ("softmax", lambda: sel.get_least_accurate_samples(PairModel(True), paths, k, mode='softmax')),
                      ("argmax", lambda: sel.get_least_accurate_samples(PairModel(True), paths, k, mode='argmax')),
                      ("unsure", lambda: sel.get_unsure_samples(PairModel(True), paths, k))):
        cap = SortedCapture()
        ref_accuracy.sorted = cap
        chosen = call()
        del ref_accuracy.sorted
        out[key + "_scores"] = np.array(cap.calls[0][0], dtype=np.float32)
        out[key + "_selected"] = paths_to_idx(chosen)

    rng = np.random.default_rng(seed + 1)
    existing = [[] if i % 2 == 0 else [(int(rng.integers(0, S - R)), int(rng.integers(0, S - R)), R, R)] for i in range(N)]
    regions, count = sel.get_least_accurate_region_maps(PairModel(True), paths, existing, R, 1)
    ex_rows = np.array([(i, *rc) for i, lst in enumerate(existing) for rc in lst], dtype=np.int64).reshape(-1, 5)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, C, S, block, R, k, batch_size], dtype=np.int64), versions=versions(),
        logits_sha=np.array(checksum(pool.logits)), unet_sha=np.array(checksum(unet)),
        existing=ex_rows, regions=regions_to_array(regions, N), count=np.int64(count), **out)
    print(f"[golden] {name}: labels={out['labels_scores'][:4]} softmax={out['softmax_scores'][:3]} regions={count}")


def adv_unet(torch, C, seed):
    """Tiny deterministic stand-in for `model.module.unet` (the accuracy predictor head): (C + 3) -> 8 -> 2 channels."""
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(C + 3, 8, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(8, 2, 3, padding=1))


def gen_accuracy_adv(name, seed, N, C, S, block, k, batch_size):
    """ActiveSelectionAccuracy.get_adversarially_vulnarable_samples (accuracy.py:73-96): gradient norm of the error head
    with respect to its input, through the reference class with a tiny convolutional `unet` (weights stored)."""
    ref = ref_shim.load_reference()
    torch = ref.torch
    from active_selection import accuracy as ref_accuracy  # reference module

    pool = make_pool(seed, N, 1, C, S, S, block)
    paths = [str(i) for i in range(N)]
    unet = adv_unet(torch, C, seed)

    class AdvModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.unet = unet

        @property
        def module(self):
            return self

        def forward(self, x):
            gs = [int(round(float(v) / ref_shim.GID_SCALE)) for v in x[:, 0, 0, 0]]
            seg = torch.from_numpy(np.stack([pool.logits[g, 0] for g in gs]))
            return seg, self.unet(torch.cat([torch.softmax(seg, dim=1), x], dim=1))

    sel = ref.active_selection.get_active_selection_class("accuracy_labels", C, pool, S, batch_size)
    cap = SortedCapture()
    ref_accuracy.sorted = cap
    chosen = sel.get_adversarially_vulnarable_samples(AdvModel(), paths, k)
    del ref_accuracy.sorted
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, C, S, block, k, batch_size], dtype=np.int64), versions=versions(),
        logits_sha=np.array(checksum(pool.logits)), scores=np.array(cap.calls[0][0], dtype=np.float32),
        selected=paths_to_idx(chosen), **{"w_" + k_.replace(".", "_"): v.detach().numpy() for k_, v in unet.state_dict().items()})
    print(f"[golden] {name}: scores={np.array(cap.calls[0][0])[:4]} selected={paths_to_idx(chosen)}")


def gen_maxsubset():
    """ActiveSelectionMaxSubset._max_representative_samples (max_subset.py:17-39): the reference's own seeded
    fixture (tests.py:616-645, seed 27, 1000 x 1024 float64, 8 candidates, 4 picks) and two float32 pools with
    duplicated candidates (np.random.randint draws with replacement in the reference's caller too)."""
    ref_shim.load_reference()
    from active_selection.max_subset import ActiveSelectionMaxSubset  # reference class

    sel = ActiveSelectionMaxSubset(None, None, None)
    np.random.seed(seed=27)
    images = np.concatenate((np.random.normal(loc=2.0, scale=1.0, size=(400, 1024)),
                             np.random.normal(loc=4.0, scale=1.0, size=(400, 1024)),
                             np.random.normal(loc=6.0, scale=1.0, size=(150, 1024)),
                             np.random.normal(loc=4.0, scale=3.0, size=(50, 1024))), axis=0)
    cand = list(np.random.randint(0, len(images), 8))
    ref_picks = sel._max_representative_samples(list(images), list(images[cand, :]), 4)
    out = {"ref_seed": np.int64(27), "ref_candidates": np.array(cand, dtype=np.int64),
           "ref_picks": np.array(ref_picks, dtype=np.int64), "ref_images_sha": np.array(checksum(images))}
    for tag, seed, N, M, D, k in (("a", 5, 600, 40, 96, 20), ("b", 6, 1500, 120, 304, 60)):
        X = synth.coreset_features(seed, N, D)
        rng = np.random.default_rng(seed)
        ci = rng.integers(0, N, size=M)                 # with replacement: duplicates happen
        Y = (X[ci] + np.float32(0.05) * rng.standard_normal((M, D)).astype(np.float32)).astype(np.float32)
        Y[M // 3] = Y[M // 5]                            # one exact duplicate pair
        picks = sel._max_representative_samples(list(X), list(Y), k)
        out[f"{tag}_meta"] = np.array([seed, N, M, D, k], dtype=np.int64)
        out[f"{tag}_picks"] = np.array(picks, dtype=np.int64)
        out[f"{tag}_x_sha"], out[f"{tag}_y_sha"] = np.array(checksum(X)), np.array(checksum(Y))
    np.savez_compressed(os.path.join(GOLDEN, "maxsubset.npz"), versions=versions(), **out)
    print(f"[golden] maxsubset: reference fixture picks {ref_picks} of candidates {cand}; a={out['a_picks'][:6]} b={out['b_picks'][:6]}")


def gen_maxsubset_poolers(name, seed, N, F_, fh, fw, crop, region_size, batch_size):
    """The three feature poolers of ActiveSelectionMaxSubset (max_subset.py:49-70, 72-86, 88-111) and
    get_representative_regions / get_representative_images on top of them, through the REFERENCE class.

    The reference pools a region crop with `F.avg_pool2d(crop, (feature_h, feature_w))` - a kernel LARGER than the crop.
    The torch the reference was written for (0.4 / 1.0) clipped such a window to the input and returned the crop mean;
    torch 2.x refuses the call.  The generator therefore runs the reference's own loops (cell enumeration, floor
    arithmetic, ordering, flattening) with ONE substitution in the reference module's namespace: an avg_pool2d whose
    kernel is clipped to the input size - the old semantics, nothing else changed."""
    ref = ref_shim.load_reference()
    torch = ref.torch
    import types
    from active_selection import max_subset as ref_ms
    real = torch.nn.functional

    def clipped_avg_pool2d(x, kernel_size, stride=None, *a, **kw):
        k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        k = (min(k[0], x.shape[-2]), min(k[1], x.shape[-1]))
        return real.avg_pool2d(x, k, stride, *a, **kw)

    ref_ms.F = types.SimpleNamespace(avg_pool2d=clipped_avg_pool2d)
    rng = np.random.Generator(np.random.Philox(key=[seed, 78]))
    feats = rng.standard_normal(size=(N, F_, fh, fw), dtype=np.float32)
    feats += (rng.integers(0, 3, size=(N, 1, 1, 1)) * 0.75).astype(np.float32)
    pool = ref_shim.SyntheticPool(np.zeros((N, 1, 2, crop, crop), np.float32), None, feats)
    sel = ref.active_selection.get_max_subset_active_selector(pool, crop, batch_size)
    paths = [str(i) for i in range(N)]
    model = lambda: ref_shim.make_replay_model(pool, "deeplab")
    cell_feats = np.stack(sel._get_features_for_image_regions(model(), paths, region_size)).astype(np.float32)
    # candidate regions: (r, c, h, w) in image coordinates, two per candidate image, one touching the border
    cand = {}
    for i in range(0, N, 2):
        cand[str(i)] = [(int(rng.integers(0, crop - region_size)), int(rng.integers(0, crop - region_size)), region_size, region_size),
                        (crop - region_size, 0, region_size, region_size)]
    li, lr = sel._convert_regions_to_list(cand)
    region_feats = np.stack(sel._get_features_for_regions(model(), li, lr)).astype(np.float32)
    selected_regions, n_sel = sel.get_representative_regions(model(), paths, cand, region_size)
    out = dict(cell_features=cell_feats, region_features=region_feats, n_selected=np.int64(n_sel),
               cand_rows=np.array([(int(k), *r) for k in sorted(cand) for r in cand[k]], dtype=np.int64),
               selected_rows=np.array([(int(k), *r) for k in selected_regions for r in selected_regions[k]], dtype=np.int64))
    if fh >= 64 and fw >= 64:       # the image-level pooler needs a 64 x 64 window (max_subset.py:79)
        out["image_features"] = np.stack(sel._get_features_for_images(model(), paths)).astype(np.float32)
        out["representative_images"] = paths_to_idx(sel.get_representative_images(model(), paths, paths[1::2]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                        meta=np.array([seed, N, F_, fh, fw, crop, region_size, batch_size], dtype=np.int64),
                        versions=versions(), features_sha=np.array(checksum(feats)), **out)
    ref_ms.F = real
    print(f"[golden] {name}: cells {cell_feats.shape}, regions {region_feats.shape}, selected {n_sel}")


def gen_noise(name, seed, N, T, C, S, block, R, k, batch_size):
    """mc_noise.py: input-noise / feature-noise / noise+dropout image scores and its region maps."""
    ref = ref_shim.load_reference()
    ref.constants.MC_STEPS = T
    pool = make_pool(seed, N, 2 * T, C, S, S, block)   # combined variants consume 2T forwards per batch
    paths = [str(i) for i in range(N)]
    sel = ref.active_selection.get_active_selection_class("noise_variance", C, pool, S, batch_size)
    cap = SortedCapture()
    ref.mc_noise.sorted = cap
    np.random.seed(0)
    s_in = sel.get_vote_entropy_for_images_with_input_noise(ref_shim.make_replay_model(pool), paths, k)
    s_ft = sel.get_vote_entropy_for_images_with_feature_noise(ref_shim.make_replay_model(pool), paths, k)
    s_cb = sel.get_vote_entropy_for_batch_with_noise_and_vote_entropy(ref_shim.make_replay_model(pool), paths, k)
    del ref.mc_noise.sorted
    existing = [[] for _ in range(N)]
    existing[1] = [(3, 5, R, R)]
    regions, count = sel.create_region_maps(ref_shim.make_replay_model(pool), paths, existing, R, 1)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, T, C, S, block, R, k, batch_size], dtype=np.int64),
        versions=versions(), logits_sha=np.array(checksum(pool.logits)),
        input_noise_scores=np.array(cap.calls[0][0], np.float32), input_noise_selected=paths_to_idx(s_in),
        feature_noise_scores=np.array(cap.calls[1][0], np.float32), feature_noise_selected=paths_to_idx(s_ft),
        combined_scores=np.array(cap.calls[2][0], np.float32), combined_selected=paths_to_idx(s_cb),
        existing=np.array([(1, 3, 5, R, R)], dtype=np.int64),
        regions=regions_to_array(regions, N), count=np.int64(count),
    )
    print(f"[golden] {name}: combined_selected={paths_to_idx(s_cb)} regions={count}")


def gen_kcenter_toy():
    """reference active_selection/tests.py:557-562."""
    ref = ref_shim.load_reference()
    sel = ref.core_set.ActiveSelectionCoreSet(None, None, None)
    feats = np.array([[1, 1], [2, 2], [2, 4], [3, 3], [4, 2], [4, 5], [5, 4], [6, 2], [7, 6]])
    picks = sel._select_batch(feats, [6], 5)
    np.savez_compressed(os.path.join(GOLDEN, "kcenter_toy.npz"), features=feats, selected=np.array([6]),
                        picks=np.array(picks, dtype=np.int64), versions=versions())
    print(f"[golden] kcenter_toy: {picks}")


def gen_coreset(name, seed, N, D, L, K):
    """_select_batch on float64 copies of float32 features (core_set.py:50,17-30)."""
    ref = ref_shim.load_reference()
    sel = ref.core_set.ActiveSelectionCoreSet(None, None, None)
    feats32 = synth.coreset_features(seed, N, D)
    picks = sel._select_batch(feats32.astype(np.float64), list(range(L)), K)
    # final min-distances, recomputed with the reference's own update rule
    md = sel._updated_distances(list(range(L)) + [int(p) for p in picks], feats32.astype(np.float64), None)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                        meta=np.array([seed, N, D, L, K], dtype=np.int64), versions=versions(),
                        features_sha=np.array(checksum(feats32)),
                        picks=np.array(picks, dtype=np.int64), min_dist=md[:, 0].astype(np.float64))
    print(f"[golden] {name}: picks[:8]={picks[:8]} max min-dist={md.max():.6f}")


def gen_coreset_e2e(name, seed, N, L, K, batch_size):
    """get_k_center_greedy_selections with the ENet geometry (128ch @64x64 -> 32x32/16 avg-pool
    -> 1152-d; core_set.py:47-49) through the replay model."""
    ref = ref_shim.load_reference()
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    feats = rng.standard_normal(size=(N, 128, 64, 64), dtype=np.float32)
    feats += (rng.integers(0, 4, size=(N, 1, 1, 1)) * 0.5).astype(np.float32)
    pool = ref_shim.SyntheticPool(np.zeros((N, 1, 2, 64, 64), np.float32), None, feats)
    sel = ref.active_selection.get_active_selection_class("coreset", 2, pool, 64, batch_size)
    already = [str(i) for i in range(L)]
    cand = [str(i) for i in range(L, N)]
    chosen = sel.get_k_center_greedy_selections(K, ref_shim.make_replay_model(pool, "enet"), cand, already)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                        meta=np.array([seed, N, L, K, batch_size], dtype=np.int64), versions=versions(),
                        features_sha=np.array(checksum(feats)), chosen=paths_to_idx(chosen))
    print(f"[golden] {name}: chosen={paths_to_idx(chosen)}")


def gen_nms_png():
    """reference active_selection/tests.py:213-231 on resources/images/nms_{0,1}.png (CPU conv2d,
    the two real maps only, min-max with -min)."""
    from PIL import Image

    ref = ref_shim.load_reference()
    torch = ref.torch
    imgs = [np.asarray(Image.open(os.path.join(ref_shim.REFERENCE_ROOT, "resources", "images", f"nms_{i}.png")))
            for i in range(2)]
    R = 127
    w = torch.ones(1, 1, R, R)
    maps = torch.stack([torch.nn.functional.conv2d(torch.from_numpy(im.astype(np.float32) / 256)[None, None], w)[0, 0]
                        for im in imgs])
    raw = maps.numpy().copy()
    mn, mx = maps.min(), maps.max()
    maps.add_(-mn).mul_(1.0 / (mx - mn))
    norm = maps.numpy().copy()
    regions, count = ref.mc_dropout.ActiveSelectionMCDropout.square_nms(maps, R, (512 * 512) // (R * R))
    rows = np.array([(i, *rc) for i, lst in enumerate(regions) for rc in lst], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "nms_png.npz"), images=np.stack(imgs).astype(np.uint8),
                        R=np.int64(R), K=np.int64((512 * 512) // (R * R)), raw_max=np.float32(raw.max()),
                        raw_maps=raw.astype(np.float32), norm_maps=norm.astype(np.float32),
                        regions=rows, count=np.int64(count), versions=versions())
    print(f"[golden] nms_png: count={count} regions={rows.tolist()}")


def _sample_pixels(seed, n, *shape):
    """n deterministic pixel coordinates inside `shape` (tuple of index arrays)."""
    rng = np.random.default_rng(seed)
    return tuple(rng.integers(0, d, size=n) for d in shape)


def gen_mc_baseline(name, seed, N, T, C, H, W, block, batch_size, n_samples=512):
    """BASELINE config-2 shape (512 x 1024, C = 19, T = 20) through the REFERENCE selectors: MC-dropout vote entropy
    (image scores + sampled map pixels) and the three single-pass CEAL scorers on pass 0.  The full maps would be
    2 MB per image, so only `n_samples` pixels per image are stored; the composed scores of all T passes (predictive
    entropy, BALD, MC confidence / margin - not in the reference) come from oracle/restate.py and are
    stored under the `composed_` prefix (restatement-pinned, not reference-pinned)."""
    from oracle import restate as R

    ref = ref_shim.load_reference()
    torch = ref.torch
    ref.constants.MC_STEPS = T
    pool = make_pool(seed, N, T, C, H, W, block)
    paths = [str(i) for i in range(N)]
    crop = H if H == W else -1
    sel = ref.active_selection.get_active_selection_class("variance", C, pool, crop, batch_size)
    cap = SortedCapture()
    ref.mc_dropout.sorted = cap
    chosen = sel.get_vote_entropy_for_images(ref_shim.make_replay_model(pool), paths, N)
    del ref.mc_dropout.sorted
    ve_scores = np.array(cap.calls[0][0], dtype=np.float32)
    ds = ref_shim.SyntheticPathsDataset(pool, paths, crop, include_labels=True)
    ve_maps = []
    for b0 in range(0, N, batch_size):
        image_batch = torch.stack([ds[i]["image"] for i in range(b0, min(N, b0 + batch_size))])
        label_batch = torch.stack([ds[i]["label"] for i in range(b0, min(N, b0 + batch_size))])
        ve_maps += [m.numpy() for m in sel._get_vote_entropy_for_batch(ref_shim.make_replay_model(pool), image_batch, label_batch)]
    ve_maps = np.stack(ve_maps).astype(np.float32)
    ceal = ref.active_selection.get_active_selection_class("ceal_entropy", C, pool, crop, batch_size)
    _, ent = ceal.get_maximum_entropy_samples(ref_shim.make_replay_model(pool), paths, N)
    cap = SortedCapture()
    ref.ceal.sorted = cap
    ceal.get_least_confident_samples(ref_shim.make_replay_model(pool), paths, N)
    ceal.get_least_margin_samples(ref_shim.make_replay_model(pool), paths, N)
    del ref.ceal.sorted
    rows, cols = _sample_pixels(seed, n_samples, H, W)
    # half of the samples on pixels that actually disagree (most pixels vote unanimously and score exactly 0)
    for i in range(N):
        nz = np.argwhere(ve_maps[i] > 0)
        if len(nz):
            pick = nz[np.random.default_rng(seed + i).integers(0, len(nz), size=n_samples // 2)]
            if i == 0:
                rows[: n_samples // 2], cols[: n_samples // 2] = pick[:, 0], pick[:, 1]
    composed = {k: [] for k in R.SCORE_NAMES}
    comp_px = {k: [] for k in ("pred_entropy", "bald", "confidence", "margin")}
    for i in range(N):
        o = R.mc_maps(pool.logits[i], pool.labels[i], C)
        sc = R.image_scores(o)
        for k in R.SCORE_NAMES:
            composed[k].append(sc[k])
        for k in comp_px:
            comp_px[k].append(o[k][rows, cols])
        assert np.array_equal(o["vote_entropy"], ve_maps[i]) or np.allclose(o["vote_entropy"], ve_maps[i], rtol=1e-6, atol=1e-7)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, T, C, H, W, block, batch_size], dtype=np.int64), versions=versions(),
        logits_sha=np.array(checksum(pool.logits)), labels_sha=np.array(checksum(pool.labels)),
        ve_scores=ve_scores, ve_selected=paths_to_idx(chosen), px_rows=rows.astype(np.int32), px_cols=cols.astype(np.int32),
        ve_px=ve_maps[:, rows, cols].astype(np.float32),
        ceal_entropy=np.array(ent, dtype=np.float32), ceal_conf=np.array(cap.calls[0][0], dtype=np.float32),
        ceal_margin=np.array(cap.calls[1][0], dtype=np.float32),
        **{"composed_" + k: np.array(v, dtype=np.float32) for k, v in composed.items()},
        **{"composed_px_" + k: np.stack(v).astype(np.float32) for k, v in comp_px.items()},
    )
    print(f"[golden] {name}: ve_scores={ve_scores} composed bald={composed['bald']}")


def gen_region_rect(name, seed, N, T, C, H, W, block, R_, selection_size, n_samples=512):
    """BASELINE config 3: rectangular 512 x 1024 planes, R = 128 (385 x 897 score maps).  The reference's
    create_region_maps is square-only (`score_maps` is allocated (base - R + 1)^2), so this fixture is
    RESTATEMENT-pinned: oracle/restate.py (itself pinned by the square reference goldens region_small / region_mid) with
    K = selection_size * H * W / R^2, the rectangular reading of mc_dropout.py:157."""
    from oracle import restate as R

    gs = list(range(N))
    logits = synth.pool_logits(seed, gs, T, C, H, W, block)
    labels = synth.pool_labels(seed, gs, H, W, C, block)
    rng = np.random.default_rng(seed + 1)
    existing = []
    for i in range(N):
        existing.append([] if i % 2 == 0 else [(int(rng.integers(0, H - R_)), int(rng.integers(0, W - R_)), R_, R_)])
    ve = [R.vote_entropy_map(R.votes_from_logits(logits[i]), C, R.valid_mask(labels[i], C)) for i in range(N)]
    maps = np.stack([R.box_sum(R.suppress_rects(m.copy(), existing[i]), R_) for i, m in enumerate(ve)])
    norm = R.minmax_normalise(maps)
    K = (selection_size * H * W) / (R_ * R_)
    regions, count = R.square_nms(norm.copy(), R_, K)
    i_, r_, c_ = _sample_pixels(seed, n_samples, N, H - R_ + 1, W - R_ + 1)
    rows = np.array([(i, *rc) for i, lst in enumerate(regions) for rc in lst], dtype=np.int64).reshape(-1, 4 + 1)
    ex_rows = np.array([(i, *rc) for i, lst in enumerate(existing) for rc in lst], dtype=np.int64).reshape(-1, 5)
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"),
        meta=np.array([seed, N, T, C, H, W, block, R_, selection_size], dtype=np.int64), versions=versions(),
        logits_sha=np.array(checksum(logits)), existing=ex_rows, regions=rows, count=np.int64(count), K=np.float64(K),
        raw_min=np.float32(maps.min()), raw_max=np.float32(maps.max()),
        px=np.stack([i_, r_, c_]).astype(np.int32), norm_px=norm[i_, r_, c_].astype(np.float32),
        pinned_by=np.array("restatement (reference is square-only)"))
    print(f"[golden] {name}: {count} regions, K={K:.1f}, raw max={maps.max():.3f}")


def gen_coreset_baseline(name, seed, N, D, L, K):
    """BASELINE config 5 size through the REFERENCE _select_batch (sklearn float64 euclidean distances)."""
    import time
    ref = ref_shim.load_reference()
    sel = ref.core_set.ActiveSelectionCoreSet(None, None, None)
    feats32 = synth.coreset_features(seed, N, D)
    t0 = time.time()
    picks = sel._select_batch(feats32.astype(np.float64), list(range(L)), K)
    dt = time.time() - t0
    md = sel._updated_distances(list(range(L)) + [int(p) for p in picks], feats32.astype(np.float64), None)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                        meta=np.array([seed, N, D, L, K], dtype=np.int64), versions=versions(),
                        features_sha=np.array(checksum(feats32)), picks=np.array(picks, dtype=np.int32),
                        min_dist_max=np.float64(md.max()), min_dist_head=md[:256, 0].astype(np.float64),
                        reference_seconds=np.float64(dt))
    print(f"[golden] {name}: picks[:8]={picks[:8]} max min-dist={md.max():.6f} ({dt:.1f} s in the reference)")


FIXTURES = {
    # odd H*W (alignment-peeling path), Pascal class count
    "mc_small": lambda: gen_mc("mc_small", synth.DEFAULT_SEED, N=6, T=5, C=21, H=65, W=65, block=8, k=3, batch_size=4),
    # H*W % 4 == 0 (128-bit path), Cityscapes class count and T, rectangular
    "mc_aligned": lambda: gen_mc("mc_aligned", synth.DEFAULT_SEED + 1, N=5, T=20, C=19, H=32, W=64, block=8, k=2, batch_size=2),
    # BASELINE config 1 shape: 16 x 513x513, C=21, T=5
    "mc_config1": lambda: gen_mc("mc_config1", synth.DEFAULT_SEED + 2, N=16, T=5, C=21, H=513, W=513, block=32, k=8, batch_size=4),
    "region_small": lambda: gen_region("region_small", synth.DEFAULT_SEED + 3, N=6, T=5, C=21, S=65, block=8, R=17, selection_size=1, batch_size=4),
    "region_mid": lambda: gen_region("region_mid", synth.DEFAULT_SEED + 4, N=5, T=8, C=19, S=129, block=16, R=33, selection_size=2, batch_size=2),
    "noise_small": lambda: gen_noise("noise_small", synth.DEFAULT_SEED + 5, N=6, T=4, C=21, S=65, block=8, R=17, k=3, batch_size=4),
    "kcenter_toy": gen_kcenter_toy,
    "coreset_small": lambda: gen_coreset("coreset_small", synth.DEFAULT_SEED + 6, N=400, D=96, L=10, K=40),
    "coreset_mid": lambda: gen_coreset("coreset_mid", synth.DEFAULT_SEED + 7, N=1500, D=2736, L=50, K=60),
    "coreset_e2e": lambda: gen_coreset_e2e("coreset_e2e", synth.DEFAULT_SEED + 8, N=14, L=4, K=5, batch_size=4),
    "nms_png": gen_nms_png,
    "maxsubset": gen_maxsubset,
    # fused final upsample: DeepLab-style exact x4 onto an odd crop (ATen's vectorised path) ...
    "upsample_odd": lambda: gen_upsample("upsample_odd", synth.DEFAULT_SEED + 9, N=5, T=5, C=21, h=17, w=17, H=65, W=65, block=2, k=2, batch_size=3),
    # ... the Cityscapes ratio (in-1)/(out-1) != 1/4 on an even, rectangular crop ...
    "upsample_rect": lambda: gen_upsample("upsample_rect", synth.DEFAULT_SEED + 10, N=4, T=4, C=19, h=12, w=16, H=48, W=64, block=2, k=2, batch_size=2),
    # ... and 33 -> 129 (3 x 3 tiles with ragged edges, 9 tiles deep windows)
    "upsample_mid": lambda: gen_upsample("upsample_mid", synth.DEFAULT_SEED + 11, N=3, T=3, C=19, h=33, w=33, H=129, W=129, block=4, k=1, batch_size=2),
    "accuracy_small": lambda: gen_accuracy("accuracy_small", 41, 6, 5, 40, 8, 9, 4, 3),
    # max-subset feature poolers: 129 x 129 feature map of a 513 crop (DeepLab geometry, 128-pixel regions -> 32 x 32 cells)
    "maxsubset_poolers": lambda: gen_maxsubset_poolers("maxsubset_poolers", 51, N=6, F_=12, fh=129, fw=129, crop=513, region_size=128, batch_size=4),
    # ... and a rectangular feature map that the crop size does not divide (floor arithmetic of max_subset.py:59-62,100-107)
    "maxsubset_poolers_rect": lambda: gen_maxsubset_poolers("maxsubset_poolers_rect", 52, N=5, F_=7, fh=33, fw=45, crop=130, region_size=40, batch_size=2),
    "accuracy_adv": lambda: gen_accuracy_adv("accuracy_adv", 43, 5, 5, 24, 8, 3, 2),
    # BASELINE sizes (VERDICT r1 "next" 1a): config 2's plane through the reference, config 3's rectangular maps through
    # the restatement, config 5's N = 10 000 through the reference's sklearn loop
    "mc_baseline": lambda: gen_mc_baseline("mc_baseline", synth.DEFAULT_SEED + 20, N=2, T=20, C=19, H=512, W=1024, block=32, batch_size=2),
    "region_rect": lambda: gen_region_rect("region_rect", synth.DEFAULT_SEED + 21, N=3, T=5, C=19, H=512, W=1024, block=32, R_=128, selection_size=2),
    "coreset_baseline": lambda: gen_coreset_baseline("coreset_baseline", synth.DEFAULT_SEED + 22, N=10000, D=2048, L=50, K=500),
}


def main(argv):
    os.makedirs(GOLDEN, exist_ok=True)
    names = argv or list(FIXTURES)
    for n in names:
        FIXTURES[n]()


if __name__ == "__main__":
    main(sys.argv[1:])
