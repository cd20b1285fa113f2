"""Mirror of reference active_selection/ceal.py (ActiveSelectionCEAL): single deterministic pass
(model.eval()), softmax -> least confidence / margin / entropy image scores -> stable ranking.
This is the T = 1 case of the Monte-Carlo kernels; one pass over the logits yields all three scores
(the reference makes a separate pass, and for the margin a D2H copy plus a host argsort, per image -
ceal.py:87-89)."""
from __future__ import annotations

import random

import torch

from .._lib import SCORE_INDEX
from .base import ActiveSelectionBase


class ActiveSelectionCEAL(ActiveSelectionBase):

    def __init__(self, dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size):
        super(ActiveSelectionCEAL, self).__init__(dataset_lmdb_env, crop_size, dataloader_batch_size)
        self.dataset_num_classes = dataset_num_classes

    def _single_pass_scores(self, model, images, weak_labels=False):
        model.eval()
        lo, hi = self._shard(images)
        chunks, weak = [], []
        for sample in self._loader(images[lo:hi], include_labels=True):
            out = self._mc_batch(model, sample['image'].cuda(), sample['label'].cuda(), 1, votes=weak_labels,
                                 probs=True, weak_labels=weak_labels)
            chunks.append(out["scores"])
            if weak_labels:
                weak.append(out["weak_labels"])
        scores = torch.cat(chunks) if chunks else torch.empty((0, len(SCORE_INDEX)), dtype=torch.float32, device="cuda")
        return scores, lo, weak

    def _column(self, scores, name, n_total):
        col = scores[:, SCORE_INDEX[name]].contiguous()
        self.last_scores = self._all_scores(col, n_total)
        return col

    def get_least_confident_samples(self, model, images, selection_count):
        scores, lo, _ = self._single_pass_scores(model, images)
        return self._rank(self._column(scores, "confidence", len(images)), lo, images, selection_count, descending=False)

    def get_least_margin_samples(self, model, images, selection_count):
        scores, lo, _ = self._single_pass_scores(model, images)
        return self._rank(self._column(scores, "margin", len(images)), lo, images, selection_count, descending=False)

    def _get_entropies(self, model, images):
        scores, _, _ = self._single_pass_scores(model, images)
        return self._all_scores(scores[:, SCORE_INDEX["pred_entropy"]].contiguous(), len(images))

    def get_maximum_entropy_samples(self, model, images, selection_count):
        scores, lo, _ = self._single_pass_scores(model, images)
        col = self._column(scores, "pred_entropy", len(images))
        return self._rank(col, lo, images, selection_count, descending=True), list(self.last_scores)

    def get_fusion_of_confidence_margin_entropy_samples(self, model, images, selection_count):
        # union of the three rankings, shuffled, first k (ceal.py:133-140) - non-deterministic by design
        scores, lo, _ = self._single_pass_scores(model, images)
        n = len(images)
        s1 = self._rank(self._column(scores, "confidence", n), lo, images, selection_count, descending=False)
        s2 = self._rank(self._column(scores, "margin", n), lo, images, selection_count, descending=False)
        s3 = self._rank(self._column(scores, "pred_entropy", n), lo, images, selection_count, descending=True)
        samples = list(set(s1 + s2 + s3))
        random.shuffle(samples)
        return samples[:selection_count]

    def get_weakly_labeled_data(self, model, images, threshold, entropies=None):
        """{path: HxW uint8 argmax labels, 255 where the ground-truth label is invalid} for every image
        whose entropy is below `threshold` (ceal.py:142-166).  Single-rank helper (no sharding)."""
        if not entropies:
            entropies = self._get_entropies(model, images)
        selected = [im for im, e in zip(images, entropies) if e < threshold]
        model.eval()
        weak = []
        for sample in self._loader(selected, include_labels=True) if selected else []:
            out = self._mc_batch(model, sample['image'].cuda(), sample['label'].cuda(), 1, votes=True, probs=False,
                                 weak_labels=True)
            weak.extend(w.cpu().numpy() for w in out["weak_labels"].unbind(0))
        return dict(zip(selected, weak))
