"""TEST INFRASTRUCTURE - runs the *unmodified* reference selectors on CPU in this container.

Only `oracle/gen_golden.py` and in-container tests use this module.  It cannot travel to
the GPU box (there is no /root/reference there); what travels are the golden vectors it
produced (tests/golden/*.npz).  Nothing on the product path imports it.

Recipe:
  1. import stubs for modules the reference imports but the scoring path never uses
     (matplotlib, lmdb, scipy.misc.imresize),
  2. CPU shims for the `torch.cuda.FloatTensor(...)` / `.cuda()` idiom
     (reference active_selection/mc_dropout.py:37, ceal.py:32),
  3. the reference's PathsDataset (dataloaders/dataset/paths_dataset.py:8-52, LMDB +
     imresize) replaced by a synthetic in-memory dataset with the same constructor,
  4. a replay model that returns pre-computed logits per forward call.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DAS_REFERENCE_ROOT", "/root/reference")
GID_SCALE = 16.0  # image[0,0,0] = GID_SCALE * g  -> survives the sigma=0.125 input noise of mc_noise.py:26


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "active_selection", "mc_dropout.py"))


_loaded = None


def load_reference():
    """Apply stubs/shims, put the reference on sys.path, return a namespace of its modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    import torch

    for name in ("matplotlib", "matplotlib.cm", "lmdb"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    import scipy.misc

    if not hasattr(scipy.misc, "imresize"):
        scipy.misc.imresize = None
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import constants  # noqa: E402  (reference module)
    from dataloaders.dataset import paths_dataset  # noqa: E402
    import active_selection  # noqa: E402
    from active_selection import mc_dropout, mc_noise, ceal, core_set  # noqa: E402

    paths_dataset.PathsDataset = SyntheticPathsDataset
    ns = types.SimpleNamespace(constants=constants, paths_dataset=paths_dataset,
                               active_selection=active_selection, mc_dropout=mc_dropout,
                               mc_noise=mc_noise, ceal=ceal, core_set=core_set, torch=torch)
    _loaded = ns
    return ns


class SyntheticPool:
    """Stands in for the LMDB env: maps a path (bytes/str of the global index) to arrays."""

    def __init__(self, logits: np.ndarray, labels: np.ndarray | None, features: np.ndarray | None = None):
        # logits [N, T_total, C, H, W]; labels [N, H, W]; features [N, F, h, w] (core-set)
        self.logits = logits
        self.labels = labels
        self.features = features

    @property
    def hw(self):
        return self.logits.shape[-2:]


class LowResPool(SyntheticPool):
    """Pool whose `logits` are the decoder's LOW-RESOLUTION outputs [N,T,C,h,w]; images / labels are H x W."""

    def __init__(self, logits, labels, H, W):
        super().__init__(logits, labels)
        self.full_hw = (H, W)

    @property
    def hw(self):
        return self.full_hw


class SyntheticPathsDataset:
    """Same constructor as the reference PathsDataset; yields {'image','label'} or the image.

    The image carries the global index g in pixel [0,0,0] (scaled by GID_SCALE) so that the
    replay model can find the logits that belong to it.
    """

    def __init__(self, env, paths, crop_size, include_labels=False):
        self.env, self.paths, self.crop_size, self.include_labels = env, paths, crop_size, include_labels

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, index):
        import torch

        g = int(self.paths[index])
        H, W = self.env.hw
        image = torch.zeros(3, H, W, dtype=torch.float32)
        image[0, 0, 0] = GID_SCALE * g
        if self.include_labels:
            return {"image": image, "label": torch.from_numpy(self.env.labels[g].astype(np.float32))}
        return image


def make_replay_model(pool: SyntheticPool, model_name: str = "deeplab"):
    """nn.Module whose k-th call on a given batch returns pass k's logits for that batch."""
    import torch

    class ReplayModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)  # something for turn_on_dropout to flip
            self.model_name = model_name
            self.return_features = False
            self.noisy_features = False
            self.calls = {}
            self.noisy_calls = 0

        @property
        def module(self):  # DataParallel-style access (reference core_set.py:44, mc_noise.py:63)
            return self

        def set_return_features(self, flag):
            self.return_features = flag

        def set_noisy_features(self, flag):
            self.noisy_features = flag

        def forward(self, x):
            gs = [int(round(float(v) / GID_SCALE)) for v in x[:, 0, 0, 0]]
            key = tuple(gs)
            t = self.calls.get(key, 0)
            self.calls[key] = t + 1
            if self.noisy_features:
                self.noisy_calls += 1
            out = torch.from_numpy(np.stack([pool.logits[g, t % pool.logits.shape[1]] for g in gs]))
            if isinstance(pool, LowResPool):
                # the reference model's own last line (models/deeplab.py:59), verbatim in meaning: the stored
                # logits play the role of `low_res_x`, the image size is the output size
                out = torch.nn.functional.interpolate(out, size=x.size()[2:], mode='bilinear', align_corners=True)
            if self.return_features:
                return out, torch.from_numpy(np.stack([pool.features[g] for g in gs]))
            return out

    return ReplayModel()
