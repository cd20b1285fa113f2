"""Shared machinery of the selector mirror (reference active_selection/base.py:1-6 + the loops that
every selector repeats).  The scoring itself happens in libdas_b200.so through ..ops."""
from __future__ import annotations

import math
import os
import sys
import types

import torch

from .. import constants as _own_constants
from .. import dist, ops
from ..prefetch import DeviceBatchLoader
from .._lib import MAX_PASS_GROUP, TOPK_MAX_K, DasError

# The reference selectors reach the data layer through the module attribute
# `paths_dataset.PathsDataset` (mc_dropout.py:131, ceal.py:21, core_set.py:42).  The data layer is out
# of scope here: by default the caller's own `dataloaders.dataset.paths_dataset` is used; tests and
# bench.py assign `paths_dataset.PathsDataset = <synthetic dataset>`.
paths_dataset = types.SimpleNamespace(PathsDataset=None)


def _dataset_class():
    if paths_dataset.PathsDataset is not None:
        return paths_dataset.PathsDataset
    try:
        from dataloaders.dataset import paths_dataset as ref_paths_dataset  # the caller's data layer
    except Exception as exc:  # pragma: no cover - depends on the embedding application
        raise DasError("no PathsDataset available: run inside the reference tree or set "
                       "deep_active_semantic_segmentation_b200.active_selection.base.paths_dataset.PathsDataset") from exc
    return ref_paths_dataset.PathsDataset


def mc_steps() -> int:
    """`constants.MC_STEPS`, read at call time like the reference does (mc_dropout.py:37,39,47)."""
    ref = sys.modules.get("constants")
    if ref is not None and hasattr(ref, "MC_STEPS"):
        return int(ref.MC_STEPS)
    return int(_own_constants.MC_STEPS)


def turn_on_dropout(model) -> None:
    """Dropout2d -> train mode so forwards are stochastic (mc_dropout.py:175-178)."""
    def _flip(m):
        if type(m) == torch.nn.Dropout2d:
            m.train()
    model.apply(_flip)


class ActiveSelectionBase:

    def __init__(self, dataset_lmdb_env, crop_size, dataloader_batch_size):
        self.crop_size = crop_size
        self.dataloader_batch_size = dataloader_batch_size
        self.env = dataset_lmdb_env
        #: Monte-Carlo passes held back and consumed by ONE launch.  1 = pure streaming (the running
        #: accumulators round-trip HBM every pass); None = as many as fit in `pass_group_bytes` of logits -
        #: with all T passes in one group the fused kernel keeps the accumulators in registers and the
        #: only HBM traffic is the logits, once ("pass groups").
        self.pass_group = None
        self.pass_group_bytes = int(os.environ.get("DAS_PASS_GROUP_BYTES", 16 << 30))
        #: scores of the last pool pass, global image order (diagnostics / parity tests)
        self.last_scores = None
        self.last_loader = None

    # -- pool iteration ------------------------------------------------------------------------
    def _shard(self, images):
        W, rank = dist.world()
        lo, hi = dist.shard_bounds(len(images), W, rank)
        return lo, hi

    def _loader(self, images, include_labels=True):
        """Batches in dataset order, as the reference's DataLoader(..., shuffle=False, num_workers=0) yields them
        (mc_dropout.py:131-132) - assembled in pinned memory by worker threads and copied asynchronously, so the
        batches arrive as CUDA tensors (prefetch.py)."""
        ds = _dataset_class()(self.env, images, self.crop_size, include_labels=include_labels)
        #: the feeder of the last pool pass: .host_seconds / .batches / .path say what the host side of the pass cost
        self.last_loader = DeviceBatchLoader(ds, self.dataloader_batch_size)
        return self.last_loader

    # -- Monte-Carlo scoring of one batch --------------------------------------------------------
    def _group_size(self, T, per_pass_bytes):
        if self.pass_group is not None:
            return max(1, min(int(self.pass_group), T))
        return max(1, min(T, MAX_PASS_GROUP, self.pass_group_bytes // max(int(per_pass_bytes), 1)))

    def _mc_batch(self, forward, image_batch, label_batch, T, votes, probs, maps=(), weak_labels=False):
        """T calls of `forward(image_batch)`; the logits are consumed in groups of G passes: all but the last
        group by K1 (das_mc_accumulate), the last group by the fused K1+K2 kernel
        (das_mc_accumulate_finalize).  G == T: nothing but the logits ever crosses HBM.

        A model that returns its LOW-RESOLUTION decoder logits (`low_res_x` of models/deeplab.py:58, i.e. the
        forward without its last line `F.interpolate(low_res_x, size=input.size()[2:], mode='bilinear',
        align_corners=True)`) is recognised by the spatial size of its output: the interpolation then happens
        inside the scoring kernel (das_mc_upsample_accumulate_finalize) and the full-resolution logits never
        exist in HBM.  Shapes the fused kernel does not take (upsampling factors below ~3.75, T > 32) are
        interpolated by torch on the device - the model's own last op - and scored by the resident-logits kernels."""
        B, _, H, W = image_batch.shape
        state, G = None, 1
        pending = []
        lowres = fuse = False
        with torch.no_grad():
            for step in range(T):
                logits = forward(image_batch)
                if state is None:
                    h, w = int(logits.shape[-2]), int(logits.shape[-1])
                    lowres = (h, w) != (H, W)
                    fuse = lowres and T <= MAX_PASS_GROUP and ops.upsample_supported(h, w, H, W)
                    G = T if fuse else self._group_size(T, B * logits.shape[1] * H * W * logits.element_size())
                    state = ops.MCState(B, logits.shape[1], H, W, T, votes=votes, probs=probs, device=logits.device,
                                        single_shot=(G >= T))
                if lowres and not fuse:
                    logits = torch.nn.functional.interpolate(logits, size=(H, W), mode='bilinear', align_corners=True)
                pending.append(logits)
                if step == T - 1:
                    if fuse:
                        return state.score_upsampled(pending, label_batch, maps=maps, scores=True,
                                                     weak_labels=weak_labels)
                    return state.score(pending, label_batch, maps=maps, scores=True, weak_labels=weak_labels)
                if len(pending) >= G:
                    state.accumulate(pending)
                    pending = []

    # -- ranking -----------------------------------------------------------------------------------
    def _rank(self, local_scores, lo, images, k, descending):
        """local_scores: f32 CUDA tensor for images[lo:lo+n].  Local K3 top-k, candidate all-gather,
        stable global merge -> tuple of the first k paths (all ranks return the same tuple)."""
        if len(images) == 0:
            raise IndexError("list index out of range")   # the reference fails on zip(*[])[1] (mc_dropout.py:195)
        k_eff = max(0, min(int(k), len(images)))
        if k_eff == 0:
            return ()
        if k_eff > TOPK_MAX_K:
            # more winners than one K3 launch ranks (the reference's sorted()[:k] has no limit): the stable sort of the
            # gathered pool scores on the host - a selection of thousands of images out of a pool is not a hot path
            vals = self._all_scores(local_scores, len(images))
            order = sorted(range(len(vals)), key=vals.__getitem__, reverse=descending)[:k_eff]
            return tuple(images[j] for j in order)
        _, ids = dist.select_ranked(local_scores, k_eff, descending, id_offset=lo)
        return tuple(images[int(j)] for j in ids)

    def _all_scores(self, local_scores, n_total):
        """Full score list in global order (list of python floats), gathered over ranks."""
        W, _ = dist.world()
        vals = local_scores.detach().cpu().tolist()
        if W == 1:
            return vals
        out = []
        for part in dist.gather_objects(vals):
            out += part
        assert len(out) == n_total
        return out


def region_tail(selector, score_maps, images, lo, region_size, selection_size):
    """Everything after the per-image box sums of create_region_maps (mc_dropout.py:152-171):
    pool min-max normalisation, image-local NMS sequences on the GPU, global order, result dict.
    score_maps is None on a rank whose shard of the pool is empty (more ranks than images): it still takes part
    in every exchange, with neutral contributions."""
    dims = None if score_maps is None else (int(score_maps.shape[1]), int(score_maps.shape[2]))
    known = [d for d in dist.gather_objects(dims) if d is not None]
    if not known:
        raise IndexError("list index out of range")      # empty pool: the reference fails on zip(*[])
    H2, W2 = known[0]
    H, W = H2 + region_size - 1, W2 + region_size - 1
    base_sq = H * W   # == base_size**2 for the reference's square crops (mc_dropout.py:129,157)
    num_requested = (selection_size * base_sq) / (region_size * region_size)
    local_mm = selector._minmax if score_maps is not None else ops.new_minmax(torch.device("cuda", torch.cuda.current_device()))
    mm = dist.allreduce_minmax(local_mm)
    if score_maps is None:
        score_maps = torch.empty((0, H2, W2), dtype=torch.float32, device=mm.device)
    else:
        ops.minmax_normalise(score_maps, mm)
    kmax = max(1, min(math.ceil(num_requested), ops.nms_pick_bound(H2, W2, region_size)))
    regions, count = global_nms(score_maps, lo, len(images), region_size, num_requested, kmax)
    new_regions = {images[i]: regions[i] for i in range(len(regions)) if regions[i]}
    return new_regions, count


def global_nms(score_maps, lo, n_images, region_size, max_selection_count, kmax):
    """The pool-global greedy NMS of mc_dropout.py:82-108 from image-local device sequences.

    Picks of one image are sorted by (score desc, flat index asc), so the reference's loop - repeat
    {global first arg-max; zero its window} - visits the candidates of all images in exactly that order: a K3
    top-k over the flattened candidate table (ties keep table order = flat order), an all-gather of each
    rank's head, and the stop rule (`count < ceil(K)`, pool max >= 0.01 checked after each pick) on the merged
    prefix.  Only the <= ceil(K) winners ever leave the device."""
    import numpy as np

    N_local, H2, W2 = score_maps.shape
    want = math.ceil(max_selection_count)
    if want > TOPK_MAX_K:       # more picks than one K3 launch ranks: k-way merge of the sequences on the host
        local = []
        if N_local > 0:
            cs, rc, cnt = ops.nms_sequences(score_maps, region_size, kmax, 0.01)
            local = dist.sequences_from_device(cs, rc, cnt)
        seqs = []
        for part in dist.gather_objects(local):
            seqs += part
        return dist.merge_nms_sequences(seqs, region_size, max_selection_count, H2, W2)
    if N_local > 0:
        cs, rc, cnt, flat = ops.nms_sequences(score_maps, region_size, kmax, 0.01, image_offset=lo, with_flat=True)
        gs, gi = dist.select_ranked(cs.reshape(-1), max(want, 1), True, ids=flat.reshape(-1))
    else:
        gs, gi = dist.select_ranked(torch.empty(0, dtype=torch.float32, device=score_maps.device), max(want, 1), True,
                                    ids=torch.empty(0, dtype=torch.int64, device=score_maps.device))
    # stop rule on the merged prefix: candidates exist (id >= 0, score > -inf), at most `want` picks, and every
    # pick after the first needs score >= 0.01 (the pool maximum the reference checks after the previous pick)
    ok = (gi >= 0) & np.isfinite(gs)
    ok[1:] &= gs[1:] >= np.float32(0.01)
    count = int(len(ok) if ok.all() else np.argmin(ok))
    img, rem = np.divmod(gi[:count], H2 * W2)
    r, c = np.divmod(rem, W2)
    regions = [[] for _ in range(n_images)]
    for i_, r_, c_ in zip(img.tolist(), r.tolist(), c.tolist()):
        regions[i_].append((r_, c_, region_size, region_size))
    return regions, count
