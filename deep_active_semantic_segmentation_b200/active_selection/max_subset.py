"""Mirror of reference active_selection/max_subset.py (ActiveSelectionMaxSubset): representativeness filter applied
after the uncertainty selection in `variance_representative` mode (active_train.py:451-452,459-460).

The greedy facility-location core (_max_representative_samples, max_subset.py:17-39) runs in libdas_b200
(das_maxsubset_greedy); the feature poolers stay PyTorch on the device (they are the tail of the network forward)
and no longer copy every pooled vector to the host one by one (max_subset.py:66-69,84-85,109-110).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from .base import ActiveSelectionBase


class ActiveSelectionMaxSubset(ActiveSelectionBase):

    def __init__(self, dataset_lmdb_env, crop_size, dataloader_batch_size):
        super(ActiveSelectionMaxSubset, self).__init__(dataset_lmdb_env, crop_size, dataloader_batch_size)

    @staticmethod
    def _as_matrix(features):
        """list of vectors / array / tensor -> contiguous CUDA [n,D] float32 or float64 (dtype preserved: sklearn
        returns float32 distances for float32 features and float64 ones for float64 features)."""
        if isinstance(features, torch.Tensor):
            t = features
        elif len(features) and isinstance(features[0], torch.Tensor):
            t = torch.stack([f.reshape(-1) for f in features])
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(features)))
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        return t.reshape(t.shape[0], -1).cuda().contiguous()

    def _max_representative_samples(self, image_features, candidate_image_features, selection_count):
        X, Y = self._as_matrix(image_features), self._as_matrix(candidate_image_features)
        if X.dtype != Y.dtype:
            X, Y = X.to(torch.float64), Y.to(torch.float64)
        print('Finding max representative candidates..')
        picks = ops.maxsubset_greedy(X, Y, selection_count).cpu().tolist()
        return [p if p >= 0 else None for p in picks]   # the reference appends None once every candidate is taken

    def _convert_regions_to_list(self, regions):
        list_images, list_regions = [], []
        for ir in sorted(list(regions.keys())):
            for r in regions[ir]:
                list_images.append(ir)
                list_regions.append(r)
        return list_images, list_regions

    def _feature_batches(self, model, images):
        model.eval()
        model.module.set_return_features(True)
        try:
            with torch.no_grad():
                for image_batch in self._loader(images, include_labels=False):
                    _, features_batch = model(image_batch.cuda())
                    yield features_batch
        finally:
            model.module.set_return_features(False)

    def _get_features_for_image_regions(self, model, images, region_size):
        """One pooled vector per non-overlapping region cell of every image, row-major over cells
        (max_subset.py:49-70; the crop mean is what avg_pool2d with a covering kernel computes)."""
        rows = []
        for fb in self._feature_batches(model, images):
            h = math.floor(region_size * fb.shape[2] / self.crop_size)
            w = math.floor(region_size * fb.shape[3] / self.crop_size)
            nr, nc = math.floor(fb.shape[2] / h), math.floor(fb.shape[3] / w)
            cells = fb[:, :, :nr * h, :nc * w].reshape(fb.shape[0], fb.shape[1], nr, h, nc, w).mean(dim=(3, 5))
            rows.append(cells.permute(0, 2, 3, 1).reshape(-1, fb.shape[1]))          # (image, row, col) order
        return torch.cat(rows).to(torch.float32)

    def _get_features_for_images(self, model, images):
        # avg_pool2d(features, (64, 64), 32) flattened channel-major (max_subset.py:72-86)
        rows = []
        for fb in self._feature_batches(model, images):
            rows.append(F.avg_pool2d(fb, (64, 64), 32).reshape(fb.shape[0], -1))
        return torch.cat(rows).to(torch.float32)

    def _get_features_for_regions(self, model, list_images, list_regions):
        # mean of the feature map over each candidate region, scaled to feature resolution (max_subset.py:88-111)
        rows, ctr = [], 0
        for fb in self._feature_batches(model, list_images):
            rr, rc = fb.shape[2] / self.crop_size, fb.shape[3] / self.crop_size
            for i in range(fb.shape[0]):
                region = list_regions[ctr + i]
                r, c = math.floor(region[0] * rr), math.floor(region[1] * rc)
                h, w = math.floor(region[2] * rr), math.floor(region[3] * rc)
                rows.append(fb[i, :, r:r + h, c:c + w].mean(dim=(1, 2)))
            ctr += fb.shape[0]
        return torch.stack(rows).to(torch.float32)

    def get_representative_regions(self, model, all_images, candidate_regions, region_size):
        candidate_list_images, candidate_list_regions = self._convert_regions_to_list(candidate_regions)
        print('Getting features for images for representativeness ..')
        all_image_features = self._get_features_for_image_regions(model, all_images, region_size)
        print('Getting features for candidates for representativeness ..')
        region_features = self._get_features_for_regions(model, candidate_list_images, candidate_list_regions)
        selected = self._max_representative_samples(all_image_features, region_features, len(region_features) // 2)
        selected_regions = {}
        for i in selected:
            selected_regions.setdefault(candidate_list_images[i], []).append(candidate_list_regions[i])
        return selected_regions, len(selected)

    def get_representative_images(self, model, all_images, candidate_images):
        print('Getting features for images for representativeness ..')
        all_image_features = self._get_features_for_images(model, all_images)
        candidate_features = self._get_features_for_images(model, candidate_images)
        selected = self._max_representative_samples(all_image_features, candidate_features, len(candidate_features) // 2)
        return [candidate_images[i] for i in selected]
