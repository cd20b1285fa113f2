"""Mirror of reference active_selection/accuracy.py (ActiveSelectionAccuracy): selection by (predicted)
segmentation error.  `model` returns the segmentation logits (get_least_accurate_sample_using_labels) or the pair
(deeplab_output, unet_output) where unet_output is the 2-channel error-predictor head (all other methods).
Scores come from one das_accuracy_scores pass per batch; the region variant reuses the vote-entropy region tail.
"""
from __future__ import annotations

import os
import time

import torch

from .. import ops
from .._lib import ACC_INDEX
from . import base
from .base import ActiveSelectionBase


class ActiveSelectionAccuracy(ActiveSelectionBase):

    def __init__(self, num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size):
        super(ActiveSelectionAccuracy, self).__init__(dataset_lmdb_env, crop_size, dataloader_batch_size)
        self.num_classes = num_classes

    def _pool_scores(self, model, images, pick_output):
        """Scores f32 [n_local, 5] of this rank's shard (device resident, no per-image sync)."""
        model.eval()
        lo, hi = self._shard(images)
        chunks = []
        with torch.no_grad():
            for sample in self._loader(images[lo:hi], include_labels=True):
                out = pick_output(model(sample['image'].cuda()))
                chunks.append(ops.accuracy_scores(out, sample['label'].cuda(), self.num_classes))
        scores = torch.cat(chunks) if chunks else torch.empty((0, len(ACC_INDEX)), dtype=torch.float32, device="cuda")
        return scores, lo

    def _select(self, scores, lo, column, images, selection_count):
        col = scores[:, ACC_INDEX[column]].contiguous()
        self.last_scores = self._all_scores(col, len(images))
        return self._rank(col, lo, images, selection_count, descending=True)

    def get_least_accurate_sample_using_labels(self, model, images, selection_count):
        # number of valid pixels whose label differs from the arg-max prediction (accuracy.py:18-37)
        scores, lo = self._pool_scores(model, images, lambda out: out)
        return self._select(scores, lo, "wrong_count", images, selection_count)

    def get_least_accurate_samples(self, model, images, selection_count, mode='softmax'):
        # predicted error mass from the 2-channel error head (accuracy.py:39-71)
        if mode not in ('softmax', 'argmax'):
            raise NotImplementedError
        scores, lo = self._pool_scores(model, images, lambda out: out[1])
        return self._select(scores, lo, "p0_sum" if mode == 'softmax' else "not_argmax_sum", images, selection_count)

    def get_adversarially_vulnarable_samples(self, model, images, selection_count):
        """Image score = mean over ALL pixels of the per-pixel L2 norm (over the input channels) of
        d sum(unet(z)) / dz at z = cat(softmax(segmentation logits), image), invalid pixels zeroed; descending
        (accuracy.py:73-96).  The forward and the backward pass through `model.module.unet` are the network's own
        (PyTorch, on the device - the reference detours through a host copy to detach z); the pool ranking is K3."""
        model.eval()
        lo, hi = self._shard(images)
        chunks = []
        for sample in self._loader(images[lo:hi], include_labels=True):
            image_batch, label_batch = sample['image'].cuda(), sample['label'].cuda()
            with torch.no_grad():
                seg_logits, _ = model(image_batch)
                z = torch.cat([torch.softmax(seg_logits, dim=1), image_batch], dim=1)
            z.requires_grad_(True)
            with torch.enable_grad():
                head = model.module.unet(z)
                head.backward(torch.ones_like(head))
            norms = torch.linalg.vector_norm(z.grad, ord=2, dim=1)                       # [B,H,W]
            norms = norms.masked_fill((label_batch < 0) | (label_batch >= self.num_classes), 0.0)
            chunks.append(norms.mean(dim=(1, 2)).to(torch.float32))
        col = torch.cat(chunks).contiguous() if chunks else torch.empty(0, dtype=torch.float32, device="cuda")
        self.last_scores = self._all_scores(col, len(images))
        return self._rank(col, lo, images, selection_count, descending=True)

    def get_unsure_samples(self, model, images, selection_count):
        # mean over valid pixels of 4 p1 - 4 p1^2 (accuracy.py:98-119)
        scores, lo = self._pool_scores(model, images, lambda out: out[1])
        selected = self._select(scores, lo, "unsure_mean", images, selection_count)
        print(self.last_scores)
        return selected

    def suppress_labeled_areas(self, score_map, labeled_region):
        """In place: zero [r:r+h, c:c+w] of a CUDA [H,W] map for every labelled (r,c,h,w) (accuracy.py:121-131)."""
        if labeled_region:
            ops.suppress_rects(score_map.unsqueeze(0), [(0, r, c, h, w) for (r, c, h, w) in labeled_region])

    def get_least_accurate_region_maps(self, model, images, existing_regions, region_size, selection_size):
        """softmax[0] of the error head (invalid pixels 0) -> labelled-region suppression -> RxR box sums ->
        pool min-max -> greedy NMS (accuracy.py:133-182); same tail as the vote-entropy region maps."""
        model.eval()
        lo, hi = self._shard(images)
        score_maps, ctr = None, 0
        self._minmax = None
        with torch.no_grad():
            for sample in self._loader(images[lo:hi], include_labels=True):
                _, unet_output = model(sample['image'].cuda())
                _, maps = ops.accuracy_scores(unet_output, sample['label'].cuda(), self.num_classes, p0_map=True)
                B, H, W = maps.shape
                if score_maps is None:
                    score_maps = torch.empty((hi - lo, H - region_size + 1, W - region_size + 1), dtype=torch.float32,
                                             device=maps.device)
                    self._minmax = ops.new_minmax(maps.device)
                rects = [(b, r, c, h, w) for b in range(B) for (r, c, h, w) in (existing_regions[lo + ctr + b] or [])]
                ops.suppress_rects(maps, rects)
                ops.box_sum(maps, region_size, self._minmax, out=score_maps[ctr:ctr + B])
                ctr += B
        out = base.region_tail(self, score_maps, images, lo, region_size, selection_size)
        model.eval()
        return out

    def wait_for_selected_samples(self, location_to_monitor, images):
        # host-only: block until an external tool has written its selection (accuracy.py:184-197)
        while not os.path.exists(location_to_monitor):
            time.sleep(5)
        with open(location_to_monitor, "r") as fptr:
            paths = [u'{}'.format(x.strip()).encode('ascii') for x in fptr.readlines() if x != '']
        return [x for x in paths if x in images]
