"""Deterministic synthetic inputs for the selection-scoring path.

Every array is a pure function of ``(seed, global image index g, pass t)`` so that any
shard of the pool, on any number of GPUs, sees identical data.

Host generators use numpy's counter-based Philox stream (used by the parity tests and
the golden generator); ``device_*`` generators build bench-sized inputs directly in HBM
with torch (same distribution, different stream - bench inputs are never compared
bit-for-bit with host ones).

Shapes follow the reference's conventions: logits are NCHW float32 as returned by the
segmentation net (reference models/deeplab.py:59), labels are float32 HxW with 255 as the
ignore value (reference dataloaders/custom_transforms.py:48-51).
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 20260
IGNORE_LABEL = 255.0


def _rng(seed: int, g: int, stream: int) -> np.random.Generator:
    # Philox takes a 2x64-bit key: (seed, image/stream id).  `stream` separates the
    # class map (0), base logits (1), labels (2), features (3) and pass t (16 + t).
    return np.random.Generator(np.random.Philox(key=[int(seed), (int(g) << 20) | int(stream)]))


def class_map(seed: int, g: int, H: int, W: int, C: int, block: int = 32) -> np.ndarray:
    """Blocky random label field: `block`x`block` pixel tiles, classes uniform in [0, C)."""
    r = _rng(seed, g, 0)
    bh, bw = -(-H // block), -(-W // block)
    coarse = r.integers(0, C, size=(bh, bw), dtype=np.int64)
    return np.repeat(np.repeat(coarse, block, axis=0), block, axis=1)[:H, :W]


def base_logits(seed: int, g: int, C: int, H: int, W: int, block: int = 32) -> np.ndarray:
    """3*onehot(class_map) + N(0,1); float32 [C,H,W]."""
    cm = class_map(seed, g, H, W, C, block)
    x = _rng(seed, g, 1).standard_normal(size=(C, H, W), dtype=np.float32)
    x[cm, np.arange(H)[:, None], np.arange(W)[None, :]] += np.float32(3.0)
    return x


def pass_logits(seed: int, g: int, t: int, C: int, H: int, W: int, block: int = 32,
                jitter: float = 0.7, base: np.ndarray | None = None) -> np.ndarray:
    """Logits of Monte-Carlo pass `t` for image `g`: base + jitter*N(0,1); float32 [C,H,W].

    Mimics dropout jitter: most pixels vote unanimously, tile borders disagree.
    """
    if base is None:
        base = base_logits(seed, g, C, H, W, block)
    n = _rng(seed, g, 16 + t).standard_normal(size=(C, H, W), dtype=np.float32)
    return (base + np.float32(jitter) * n).astype(np.float32)


def labels(seed: int, g: int, H: int, W: int, C: int, block: int = 32, border: int = 16,
           ignore_frac: float = 0.05) -> np.ndarray:
    """float32 [H,W]: the class map with a `border`-px frame and `ignore_frac` random pixels = 255."""
    lab = class_map(seed, g, H, W, C, block).astype(np.float32)
    r = _rng(seed, g, 2)
    lab[r.random(size=(H, W)) < ignore_frac] = IGNORE_LABEL
    b = min(border, H // 4, W // 4)
    if b > 0:
        lab[:b, :] = IGNORE_LABEL
        lab[-b:, :] = IGNORE_LABEL
        lab[:, :b] = IGNORE_LABEL
        lab[:, -b:] = IGNORE_LABEL
    return lab


def pool_logits(seed: int, gs, T: int, C: int, H: int, W: int, block: int = 32) -> np.ndarray:
    """float32 [len(gs), T, C, H, W] for a list of global image indices."""
    out = np.empty((len(gs), T, C, H, W), dtype=np.float32)
    for i, g in enumerate(gs):
        b = base_logits(seed, g, C, H, W, block)
        for t in range(T):
            out[i, t] = pass_logits(seed, g, t, C, H, W, block, base=b)
    return out


def pool_labels(seed: int, gs, H: int, W: int, C: int, block: int = 32) -> np.ndarray:
    return np.stack([labels(seed, g, H, W, C, block) for g in gs]).astype(np.float32)


def coreset_features(seed: int, N: int, D: int, n_clusters: int = 8) -> np.ndarray:
    """float32 [N,D]: N(0,1) scaled by (1 + cluster offset) around `n_clusters` centres."""
    r = _rng(seed, 0, 3)
    centres = r.standard_normal(size=(n_clusters, D), dtype=np.float32) * np.float32(2.0)
    assign = r.integers(0, n_clusters, size=N)
    scale = (1.0 + 0.25 * assign).astype(np.float32)[:, None]
    x = r.standard_normal(size=(N, D), dtype=np.float32) * scale + centres[assign]
    return np.ascontiguousarray(x, dtype=np.float32)


# --------------------------------------------------------------------------------------
# device-side generators (bench sizes; torch is imported lazily so the oracle can use the
# numpy half of this module without torch)
# --------------------------------------------------------------------------------------

def device_pass_logits(seed: int, first_g: int, B: int, T: int, C: int, H: int, W: int,
                       device, block: int = 32, jitter: float = 0.7):
    """List of T float32 CUDA tensors [B,C,H,W] (one per MC pass) + labels [B,H,W].

    Same construction as the host generators (blocky class map, +3 on the map's class,
    unit noise, per-pass jitter); the random stream is torch's Philox seeded from
    (seed, first_g), so a given (seed, first_g, B) is reproducible on any rank.
    """
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed((int(seed) << 24) ^ (int(first_g) * 2654435761 % (1 << 31)))
    bh, bw = -(-H // block), -(-W // block)
    coarse = torch.randint(0, C, (B, bh, bw), generator=gen, device=device)
    cm = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].contiguous()
    base = torch.randn((B, C, H, W), generator=gen, device=device, dtype=torch.float32)
    base.scatter_add_(1, cm[:, None], torch.full((B, 1, H, W), 3.0, device=device))
    passes = []
    for _ in range(T):
        n = torch.randn((B, C, H, W), generator=gen, device=device, dtype=torch.float32)
        passes.append(n.mul_(jitter).add_(base))
    lab = cm.to(torch.float32)
    ign = torch.rand((B, H, W), generator=gen, device=device) < 0.05
    lab[ign] = IGNORE_LABEL
    b = min(16, H // 4, W // 4)
    if b > 0:
        lab[:, :b, :] = IGNORE_LABEL
        lab[:, -b:, :] = IGNORE_LABEL
        lab[:, :, :b] = IGNORE_LABEL
        lab[:, :, -b:] = IGNORE_LABEL
    return passes, lab.contiguous()
