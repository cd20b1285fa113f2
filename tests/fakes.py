"""Synthetic data layer + replay model for driving the selector mirror in tests / smoke / bench.

Same idea as oracle/ref_shim.py (which drives the *reference* classes): the dataset yields an image
that carries its global index, the model replays pre-computed logits pass by pass.  Here everything
lives on the GPU and nothing of the reference is imported.
"""
import numpy as np
import torch

GID_SCALE = 16.0


class Pool:
    def __init__(self, logits, labels, features=None):
        # logits [N,T_total,C,H,W] float32 (numpy), labels [N,H,W] float32 or None, features [N,F,h,w]
        self.logits, self.labels, self.features = logits, labels, features

    @property
    def hw(self):
        return self.logits.shape[-2:]


class LowResPool(Pool):
    """`logits` are the decoder's low-resolution outputs [N,T,C,h,w] (models/deeplab.py:58 `low_res_x`);
    images and labels are H x W.  A ReplayModel on this pool returns the low-resolution logits as they are:
    it plays a model whose forward stops before the final F.interpolate."""

    def __init__(self, logits, labels, H, W):
        super().__init__(logits, labels)
        self.full_hw = (H, W)

    @property
    def hw(self):
        return self.full_hw


class SyntheticPathsDataset(torch.utils.data.Dataset):
    def __init__(self, env, paths, crop_size, include_labels=False):
        self.env, self.paths, self.crop_size, self.include_labels = env, paths, crop_size, include_labels

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, index):
        g = int(self.paths[index])
        H, W = self.env.hw
        image = torch.zeros(3, H, W, dtype=torch.float32)
        image[0, 0, 0] = GID_SCALE * g
        if self.include_labels:
            return {"image": image, "label": torch.from_numpy(self.env.labels[g].astype(np.float32))}
        return image


class ReplayModel(torch.nn.Module):
    """k-th call on a given batch returns pass k of that batch's logits (uploaded to the GPU)."""

    def __init__(self, pool, model_name="deeplab", device="cuda"):
        super().__init__()
        self.drop = torch.nn.Dropout2d(0.25)
        self.pool, self.model_name, self.dev = pool, model_name, device
        self.return_features = False
        self.noisy_features = False
        self.calls = {}
        self.noisy_calls = 0
        self.dropout_train_calls = 0

    @property
    def module(self):
        return self

    def set_return_features(self, flag):
        self.return_features = flag

    def set_noisy_features(self, flag):
        self.noisy_features = flag

    def forward(self, x):
        gs = [int(round(float(v) / GID_SCALE)) for v in x[:, 0, 0, 0].cpu()]
        key = tuple(gs)
        t = self.calls.get(key, 0)
        self.calls[key] = t + 1
        self.noisy_calls += int(self.noisy_features)
        self.dropout_train_calls += int(self.drop.training)
        T_total = self.pool.logits.shape[1]
        out = torch.from_numpy(np.stack([self.pool.logits[g, t % T_total] for g in gs])).to(self.dev)
        if self.return_features:
            return out, torch.from_numpy(np.stack([self.pool.features[g] for g in gs])).to(self.dev)
        return out
