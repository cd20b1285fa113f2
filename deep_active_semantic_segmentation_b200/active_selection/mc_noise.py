"""Mirror of reference active_selection/mc_noise.py (ActiveSelectionMCNoise).

Same vote-entropy reduction as MC-dropout; only the source of stochasticity differs:
input noise N(0, 0.125) (mc_noise.py:26), the model's own feature noise (`set_noisy_features`,
mc_noise.py:63,83) or dropout (mc_noise.py:88-91); the combined score adds two entropy maps
per pixel (mc_noise.py:141-143,165-167).  The input noise is drawn on the device, one `das_add_gaussian_noise` launch
per pass (the reference draws it with numpy on the host and uploads it every pass - SURVEY.md section 8(f) item 4).
"""
from __future__ import annotations

import torch

from .. import ops
from .._lib import SCORE_INDEX
from .base import mc_steps, turn_on_dropout
from .mc_dropout import ActiveSelectionMCDropout

INPUT_NOISE_SIGMA = 0.125  # mc_noise.py:26

_VE = SCORE_INDEX["vote_entropy"]


class ActiveSelectionMCNoise(ActiveSelectionMCDropout):

    def __init__(self, num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size):
        super(ActiveSelectionMCNoise, self).__init__(num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size)
        self._noise_passes = 0      # stream id of the next input-noise draw (one independent Philox stream per pass)

    def _noisy_forward(self, model):
        """x -> model(x + N(0, 0.125)) with fresh noise per call (mc_noise.py:26-27), drawn by das_add_gaussian_noise:
        one HBM-bound pass instead of torch's generator + add (28 -> 8 us per pass for eight 513 x 513 images, which was as
        much as the scoring kernel).  The stream is keyed by torch's seed (torch.manual_seed controls it) and a
        per-selector pass counter, so every pass of every batch gets its own independent noise."""
        from .. import dist
        seed = int(torch.initial_seed())
        rank = dist.world()[1]          # ranks score different images: different streams, not the same noise on each
        buf = {}

        def forward(x):
            self._noise_passes += 1
            key = (tuple(x.shape), x.device)
            if key not in buf:          # one scratch tensor per batch shape: the model consumes it before the next pass
                buf.clear()
                buf[key] = torch.empty_like(x)
            return model(ops.add_gaussian_noise(x, INPUT_NOISE_SIGMA, seed, (rank << 40) | self._noise_passes, out=buf[key]))

        return forward

    # ---- per-batch maps (lists of [H,W] CUDA tensors, as in the reference) -----------------------
    def _ve(self, forward, image_batch, label_batch, maps=True):
        return self._mc_batch(forward, image_batch, label_batch, mc_steps(), votes=True, probs=False,
                              maps=("vote_entropy",) if maps else ())

    def _get_vote_entropy_for_batch_with_input_noise(self, model, image_batch, label_batch):
        noisy_forward = self._noisy_forward(model)
        return list(self._ve(noisy_forward, image_batch, label_batch)["vote_entropy"].unbind(0))

    def _feature_noise(self, model, image_batch, label_batch, maps=True):
        model.module.set_noisy_features(True)
        try:
            return self._ve(model, image_batch, label_batch, maps)
        finally:
            model.module.set_noisy_features(False)

    def _get_vote_entropy_for_batch_with_feature_noise(self, model, image_batch, label_batch):
        return list(self._feature_noise(model, image_batch, label_batch)["vote_entropy"].unbind(0))

    def _dropout(self, model, image_batch, label_batch, maps=True):
        turn_on_dropout(model)
        try:
            return self._ve(model, image_batch, label_batch, maps)
        finally:
            model.eval()

    def _get_vote_entropy_for_batch_with_mc_dropout(self, model, image_batch, label_batch):
        return list(self._dropout(model, image_batch, label_batch)["vote_entropy"].unbind(0))

    # ---- pool scoring ----------------------------------------------------------------------------
    def _pool(self, images, batch_scores_fn):
        lo, hi = self._shard(images)
        chunks = []
        for sample in self._loader(images[lo:hi], include_labels=True):
            chunks.append(batch_scores_fn(sample['image'].cuda(), sample['label'].cuda()))
        col = torch.cat(chunks) if chunks else torch.empty(0, dtype=torch.float32, device="cuda")
        self.last_scores = self._all_scores(col, len(images))
        return col.contiguous(), lo

    def get_vote_entropy_for_images_with_input_noise(self, model, images, selection_count):
        model.eval()

        noisy_forward = self._noisy_forward(model)

        col, lo = self._pool(images, lambda x, y: self._ve(noisy_forward, x, y, maps=False)["scores"][:, _VE])
        return self._rank(col, lo, images, selection_count, descending=True)

    # ---- BASELINE config 4 (composed): CEAL entropy / margin / confidence over T input-noise passes ----
    def get_mc_scores_for_images_with_input_noise(self, model, images, selection_count, score="pred_entropy"):
        """The perturbation of mc_noise.py:26-27 (every pass sees image + N(0, 0.125), model in eval mode) with the
        scores of ceal.py evaluated on the Monte-Carlo mean softmax: 'pred_entropy' (ceal.py:111-123, descending),
        'margin' / 'confidence' (ceal.py:84-97, 36-39; ascending), plus 'bald' / 'expected_entropy' / 'vote_entropy'
        from the same pass over the logits.  With constants.MC_STEPS = 1 and sigma -> 0 this is exactly CEAL.
        Returns (paths, {score name: [scores of the whole pool]}), like get_mc_scores_for_images."""
        if score not in SCORE_INDEX:
            raise NotImplementedError(score)
        model.eval()

        noisy_forward = self._noisy_forward(model)

        scores, lo = self._pool_scores(model, images, mc_steps(), votes=True, probs=True, forward=noisy_forward)
        allv = {name: self._all_scores(scores[:, j].contiguous(), len(images)) for name, j in SCORE_INDEX.items()}
        self.last_scores = allv[score]
        descending = score not in ("confidence", "margin")
        return self._rank(scores[:, SCORE_INDEX[score]].contiguous(), lo, images, selection_count, descending), allv

    def get_vote_entropy_for_images_with_feature_noise(self, model, images, selection_count):
        model.eval()
        col, lo = self._pool(images, lambda x, y: self._feature_noise(model, x, y, maps=False)["scores"][:, _VE])
        return self._rank(col, lo, images, selection_count, descending=True)

    def get_vote_entropy_for_batch_with_noise_and_vote_entropy(self, model, images, selection_count):
        model.eval()

        def both(x, y):
            # mean(a + b) over the image == mean(a) + mean(b); the per-pixel sum (mc_noise.py:141) is only
            # materialised on the region path below
            a = self._feature_noise(model, x, y, maps=False)["scores"][:, _VE]
            b = self._dropout(model, x, y, maps=False)["scores"][:, _VE]
            return a + b

        col, lo = self._pool(images, both)
        return self._rank(col, lo, images, selection_count, descending=True)

    def create_region_maps(self, model, images, existing_regions, region_size, selection_size):
        def batch_maps(image_batch, label_batch):
            a = self._feature_noise(model, image_batch, label_batch)["vote_entropy"]
            b = self._dropout(model, image_batch, label_batch)["vote_entropy"]
            ops.add_maps(a, b)   # combined_entropies = x + y (mc_noise.py:165)
            return a

        out = self._region_maps_from(batch_maps, images, existing_regions, region_size, selection_size)
        model.eval()
        return out
