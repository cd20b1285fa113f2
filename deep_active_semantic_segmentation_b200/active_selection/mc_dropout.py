"""Mirror of reference active_selection/mc_dropout.py (ActiveSelectionMCDropout): same class, same
method signatures and return types; the scoring bodies run in libdas_b200.so.

Reference-pinned outputs: per-pixel vote entropy over T argmax votes (mc_dropout.py:30-80), image
mean + stable descending top-k (:173-196), sliding-region sums + pool min-max + greedy NMS (:82-171).
North-star additions (composed): get_mc_scores_for_images -> predictive entropy / BALD /
confidence / margin of the MC-mean softmax in the same pass over the logits.
"""
from __future__ import annotations

import math
import random

import torch

from .. import ops
from .._lib import SCORE_INDEX
from . import base
from .base import ActiveSelectionBase, mc_steps, turn_on_dropout


class ActiveSelectionMCDropout(ActiveSelectionBase):

    def __init__(self, dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size):
        super(ActiveSelectionMCDropout, self).__init__(dataset_lmdb_env, crop_size, dataloader_batch_size)
        self.dataset_num_classes = dataset_num_classes

    def get_random_uncertainity(self, images, selection_count):
        # host-only baseline, no scoring involved (mc_dropout.py:23-28)
        scores = [random.random() for _ in images]
        order = sorted(range(len(images)), key=lambda j: scores[j], reverse=True)
        return tuple(images[j] for j in order)[:selection_count]

    def _get_vote_entropy_for_batch(self, model, image_batch, label_batch):
        """-> list of B vote-entropy maps [H,W] (CUDA float32), as mc_dropout.py:30-80."""
        out = self._mc_batch(model, image_batch, label_batch, mc_steps(), votes=True, probs=False,
                             maps=("vote_entropy",))
        return list(out["vote_entropy"].unbind(0))

    @staticmethod
    def square_nms(score_maps, region_size, max_selection_count):
        """Same contract as mc_dropout.py:82-108: score_maps is a float32 [N,H',W'] tensor (CPU in the
        reference; CPU or CUDA here) and is MUTATED; returns (list of per-image [(r,c,R,R)], count).
        The image-local pick sequences are produced by the CUDA NMS kernel, the global order by the
        (score desc, flat index asc) merge with the reference's stop rule."""
        N, H2, W2 = score_maps.shape
        dev_maps = score_maps if score_maps.is_cuda else score_maps.cuda()
        dev_maps = dev_maps.contiguous()
        if dev_maps.data_ptr() == score_maps.data_ptr():
            # the NMS kernel zeroes the windows of EVERY image-local pick; the reference's loop only those of the picks
            # it actually takes (mc_dropout.py:97-103): work on a copy, replay the chosen windows on the caller's tensor
            dev_maps = dev_maps.clone()
        kmax = max(1, min(math.ceil(max_selection_count), ops.nms_pick_bound(H2, W2, region_size)))
        regions, count = base.global_nms(dev_maps, 0, N, region_size, max_selection_count, kmax)
        # leave the caller's tensor in the state the reference leaves it in: chosen windows zeroed
        for i, lst in enumerate(regions):
            for (r, c, _, _) in lst:
                score_maps[i, max(0, r - region_size):min(H2, r + region_size),
                           max(0, c - region_size):min(W2, c + region_size)] = 0
        return regions, count

    @staticmethod
    def suppress_labeled_entropy(entropy_map, labeled_region):
        """In place: zero [r:r+h, c:c+w] of a CUDA [H,W] map for every labelled (r,c,h,w) (mc_dropout.py:110-121)."""
        if labeled_region:
            ops.suppress_rects(entropy_map.unsqueeze(0), [(0, r, c, h, w) for (r, c, h, w) in labeled_region])

    def _region_maps_from(self, batch_maps_fn, images, existing_regions, region_size, selection_size):
        """Shared by mc_dropout / mc_noise create_region_maps: batch_maps_fn(image_batch, label_batch) ->
        f32 [B,H,W] uncertainty maps; rest is suppression -> box sum -> min-max -> NMS."""
        lo, hi = self._shard(images)
        score_maps = None
        ctr = 0
        self._minmax = None
        for sample in self._loader(images[lo:hi], include_labels=True):
            image_batch = sample['image'].cuda()
            label_batch = sample['label'].cuda()
            maps = batch_maps_fn(image_batch, label_batch)
            B, H, W = maps.shape
            if score_maps is None:
                score_maps = torch.empty((hi - lo, H - region_size + 1, W - region_size + 1), dtype=torch.float32,
                                         device=maps.device)
                self._minmax = ops.new_minmax(maps.device)
            rects = [(b, r, c, h, w) for b in range(B) for (r, c, h, w) in (existing_regions[lo + ctr + b] or [])]
            ops.suppress_rects(maps, rects)
            ops.box_sum(maps, region_size, self._minmax, out=score_maps[ctr:ctr + B])
            ctr += B
        return base.region_tail(self, score_maps, images, lo, region_size, selection_size)

    def create_region_maps(self, model, images, existing_regions, region_size, selection_size):
        turn_on_dropout(model)
        T = mc_steps()

        def batch_maps(image_batch, label_batch):
            return self._mc_batch(model, image_batch, label_batch, T, votes=True, probs=False,
                                  maps=("vote_entropy",))["vote_entropy"]

        out = self._region_maps_from(batch_maps, images, existing_regions, region_size, selection_size)
        model.eval()
        return out

    def _pool_scores(self, model, images, T, votes, probs, forward=None):
        """Scores [n_local, 6] of this rank's shard, kept on the device (no per-image sync)."""
        lo, hi = self._shard(images)
        chunks = []
        fwd = model if forward is None else forward
        for sample in self._loader(images[lo:hi], include_labels=True):
            image_batch = sample['image'].cuda()
            label_batch = sample['label'].cuda()
            chunks.append(self._mc_batch(fwd, image_batch, label_batch, T, votes, probs)["scores"])
        if chunks:
            return torch.cat(chunks), lo
        return torch.empty((0, len(SCORE_INDEX)), dtype=torch.float32, device="cuda"), lo

    def get_vote_entropy_for_images(self, model, images, selection_count):
        turn_on_dropout(model)
        scores, lo = self._pool_scores(model, images, mc_steps(), votes=True, probs=False)
        model.eval()
        col = scores[:, SCORE_INDEX["vote_entropy"]].contiguous()
        self.last_scores = self._all_scores(col, len(images))
        return self._rank(col, lo, images, selection_count, descending=True)

    # ---- north-star addition (not in the reference): softmax-mean entropy / BALD in the same pass ----
    def get_mc_scores_for_images(self, model, images, selection_count, score="bald"):
        """Top-k by a composed MC score ('pred_entropy', 'bald', 'expected_entropy' descending;
        'confidence', 'margin' ascending, as in ceal.py:69,97).  Returns (paths, {name: [scores]})."""
        turn_on_dropout(model)
        scores, lo = self._pool_scores(model, images, mc_steps(), votes=True, probs=True)
        model.eval()
        allv = {name: self._all_scores(scores[:, j].contiguous(), len(images)) for name, j in SCORE_INDEX.items()}
        self.last_scores = allv[score]
        descending = score not in ("confidence", "margin")
        return self._rank(scores[:, SCORE_INDEX[score]].contiguous(), lo, images, selection_count, descending), allv
