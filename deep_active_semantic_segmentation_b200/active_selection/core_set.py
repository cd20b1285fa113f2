"""Mirror of reference active_selection/core_set.py (ActiveSelectionCoreSet): k-center greedy over
average-pooled decoder features.  Feature extraction stays in PyTorch on the device (no per-image
D2H, core_set.py:63); distances / min-update / arg-max run in the K4 kernels (fp64 accumulation)."""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn.functional as F

from .. import dist, ops
from .._lib import DasError
from .base import ActiveSelectionBase

_POOL = {"deeplab": (64, 64), "enet": (32, 32)}  # core_set.py:44-49


class ActiveSelectionCoreSet(ActiveSelectionBase):

    def __init__(self, dataset_lmdb_env, crop_size, dataloader_batch_size):
        super(ActiveSelectionCoreSet, self).__init__(dataset_lmdb_env, crop_size, dataloader_batch_size)
        self.last_min_distances = None
        #: tensor-core (tcgen05 bf16) distance filter in front of the exact float64 update: "auto" = when the
        #: feature matrix is large enough to matter and the N x rows table fits in free HBM; True / False force it.
        #: Selections are bit-identical either way.
        self.tensor_core_filter = "auto"
        self.last_filter_stats = None
        #: multi-GPU: "auto" = every rank runs the whole (deterministic) greedy loop on its own GPU while the pool
        #: fits the single-cluster kernel (N <= 32768): 500 latency-bound steps cost ~4 ms locally but a collective
        #: per step (~60 us) when the rows of min_d are sharded.  True forces row sharding (larger pools, tests).
        self.shard_rows = "auto"
        self.replicated_rows_limit = 32768
        self._budget_ok = {}
        self.last_forward_rows = None

    def _make_filter(self, feats, lo, hi):
        n, d = feats.shape
        use = self.tensor_core_filter
        if use == "auto":
            use = n * d >= (1 << 20)
            if use:
                # cudaMemGetInfo costs ~1 ms (a quarter of the whole selection at N = 10 000): asked once per shape
                key = (n, d, hi - lo, str(feats.device))
                if key not in self._budget_ok:
                    self._budget_ok[key] = ops.kcenter_filter_budget_ok(n, d, hi - lo, feats.device)
                use = self._budget_ok[key]
        return ops.KCenterFilter(feats, lo, hi) if use else None

    @staticmethod
    def _as_device_features(features):
        """[n, D] numpy / tensor of any float type -> float32 CUDA tensor.  The reference's matrix is float64 but holds
        float32 network outputs (core_set.py:50,63), which float32 carries exactly; genuinely wider values are rounded
        to float32 with a warning (the kernels read float32 rows and accumulate in float64)."""
        if isinstance(features, torch.Tensor):
            f = features
            if f.dtype != torch.float32:
                f32 = f.to(torch.float32)
                if not torch.equal(f32.to(f.dtype), f):
                    warnings.warn("core-set features are not exactly representable in float32: rounded to float32 "
                                  "(the reference stores float32 network outputs, core_set.py:50,63)", RuntimeWarning)
                f = f32
            return f.cuda().contiguous()
        a = np.asarray(features)
        a32 = a.astype(np.float32)
        if not np.array_equal(a32.astype(np.float64), a.astype(np.float64)):
            warnings.warn("core-set features are not exactly representable in float32: rounded to float32 "
                          "(the reference stores float32 network outputs, core_set.py:50,63)", RuntimeWarning)
        return torch.from_numpy(np.ascontiguousarray(a32)).cuda()

    def _updated_distances(self, cluster_centers, features, min_distances):
        """Same contract as core_set.py:32-38: euclidean distances (float64) of every row of `features` to the rows
        `cluster_centers`; min over the centres as an [n, 1] column when `min_distances` is None, otherwise
        np.minimum(min_distances, distances) (an [n, len(cluster_centers)] array, [n, 1] for the reference's only call
        shape `[ind]`).  The distances come from the K4 init kernel (exact float32 differences accumulated in float64;
        sklearn's expanded form differs from that in the last few ulp).  Returns numpy, like the reference."""
        feats = self._as_device_features(features)
        n = feats.shape[0]
        centres = [int(c) for c in cluster_centers]
        if len(centres) == 0:
            raise ValueError("Found array with 0 sample(s)")       # what sklearn.pairwise_distances raises
        dev = feats.device
        key = torch.zeros(2, dtype=torch.int64, device=dev)

        def column(cs):
            d2 = torch.empty(n, dtype=torch.float64, device=dev)
            ops.kcenter_init(feats, 0, n, torch.as_tensor(cs, dtype=torch.int32, device=dev), d2, key)
            return d2.sqrt_()

        if min_distances is None:
            return column(centres).cpu().numpy().reshape(-1, 1)
        dist_cols = torch.stack([column([c]) for c in centres], dim=1).cpu().numpy()
        return np.minimum(np.asarray(min_distances), dist_cols)

    def _select_batch(self, features, selected_indices, N):
        """features [n,D] (numpy / tensor), selected_indices list[int], N picks -> list[int] (core_set.py:17-30)."""
        feats = self._as_device_features(features)
        selected = [int(i) for i in selected_indices]
        if len(selected) == 0:
            raise ValueError("k-center needs at least one already selected row (core_set.py:19)")
        W, rank = dist.world()
        shard = self.shard_rows
        if shard == "auto":
            shard = feats.shape[0] > self.replicated_rows_limit
        if W == 1 or not shard:
            flt = self._make_filter(feats, 0, feats.shape[0])
            picks, min_d = ops.kcenter_greedy(feats, selected, N, flt)
            picks = picks.cpu().tolist()
            self.last_filter_stats = flt.stats() if flt is not None else None
        else:
            picks, min_d = self._select_batch_sharded(feats, selected, N, W, rank)
        self.last_min_distances = min_d
        for ind in picks:
            # "New examples should not be in already selected" (core_set.py:23-25)
            assert ind not in selected
        if min_d.numel():
            print('Maximum distance from cluster centers is %0.5f' % float(min_d.max()))
        return picks

    def _select_batch_sharded(self, feats, selected, N, W, rank):
        """Rows of min_d sharded over ranks, features replicated; one 16-byte exchange per step
        (value bits, row index) - the arg-max all-reduce of SURVEY.md section 8(e)(iii)."""
        import torch.distributed as td

        n = feats.shape[0]
        lo, hi = dist.shard_bounds(n, W, rank)
        if hi <= lo:
            raise DasError("this rank received an empty shard of the feature rows")
        dev = feats.device
        min_d2 = torch.empty(hi - lo, dtype=torch.float64, device=dev)
        key = torch.zeros(2, dtype=torch.int64, device=dev)
        centre = torch.zeros(1, dtype=torch.int32, device=dev)
        cen = torch.as_tensor(selected, dtype=torch.int32, device=dev)
        flt = self._make_filter(feats, lo, hi)
        ops.kcenter_init(feats, lo, hi, cen, min_d2, key, flt)
        picks = []
        keys = [torch.empty_like(key) for _ in range(W)]
        for _ in range(N):
            td.all_gather(keys, key)
            allk = torch.stack(keys)                       # [W,2]: (fp64 bits of the shard max, row)
            best = allk[:, 0].max()
            row = torch.where(allk[:, 0] == best, allk[:, 1], torch.full_like(allk[:, 1], 2 ** 62)).min()
            centre.copy_(row.to(torch.int32).reshape(1))
            picks.append(row)
            ops.kcenter_step(feats, lo, hi, centre, min_d2, key, flt)
        parts = [torch.empty(dist.shard_bounds(n, W, r)[1] - dist.shard_bounds(n, W, r)[0], dtype=torch.float64, device=dev)
                 for r in range(W)]
        td.all_gather(parts, min_d2)
        self.last_filter_stats = flt.stats() if flt is not None else None
        return [int(p) for p in torch.stack(picks).cpu().tolist()] if picks else [], torch.cat(parts).sqrt()

    def _pooled_features(self, model, combined_paths):
        """[len(combined_paths), D] float32 on the device.  The forwards - the real cost of this selector - are sharded
        by image over the ranks (each rank runs the network on its contiguous slice only) and the pooled rows are
        all-gathered; the greedy loop that follows is replicated (see `shard_rows`)."""
        name = model.module.model_name
        if name not in _POOL:
            raise NotImplementedError(name)
        ks = _POOL[name]
        lo, hi = self._shard(combined_paths)
        rows = []
        model.eval()
        model.module.set_return_features(True)
        try:
            with torch.no_grad():
                for sample in self._loader(combined_paths[lo:hi], include_labels=False):
                    _, fb = model(sample.cuda())
                    fb = F.avg_pool2d(fb, ks, ks[0] // 2)
                    rows.append(fb.reshape(fb.shape[0], -1).to(torch.float32))   # channel-major flatten (core_set.py:63)
        finally:
            model.module.set_return_features(False)
        local = torch.cat(rows).contiguous() if rows else torch.empty((0, 0), dtype=torch.float32, device="cuda")
        self.last_forward_rows = hi - lo
        return dist.all_gather_rows(local, len(combined_paths))

    def get_k_center_greedy_selections(self, selection_size, model, candidate_image_batch, already_selected_image_batch):
        combined_paths = already_selected_image_batch + candidate_image_batch
        features = self._pooled_features(model, combined_paths)
        selected_indices = self._select_batch(features, list(range(len(already_selected_image_batch))), selection_size)
        return [combined_paths[i] for i in selected_indices]
