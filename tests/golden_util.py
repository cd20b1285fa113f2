"""Shared helpers for the golden-vector tests (inputs are regenerated from the stored seeds)."""
import hashlib
import os

import numpy as np

from deep_active_semantic_segmentation_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def mc_inputs(g, n=None, t_total=None):
    """Regenerate (logits [N,T,C,H,W], labels [N,H,W]) for an mc_* / region_* / noise_* fixture."""
    meta = [int(v) for v in g["meta"]]
    return meta


def pool_from_meta(seed, N, T, C, H, W, block, expect_sha=None, n=None):
    gs = list(range(N if n is None else n))
    logits = synth.pool_logits(seed, gs, T, C, H, W, block)
    labels = synth.pool_labels(seed, gs, H, W, C, block)
    if expect_sha is not None and n is None:
        assert sha(logits) == str(expect_sha), "synthetic input stream changed - regenerate goldens"
    return logits, labels


def regions_from_rows(rows, N):
    out = [[] for _ in range(N)]
    for i, r, c, h, w in rows.tolist():
        out[i].append((r, c, h, w))
    return out


def upsample_case(name):
    """-> (golden, meta dict, low-res logits [N,T,C,h,w], labels [N,H,W]) of an upsample_* fixture."""
    g = load(name)
    keys = ("seed", "N", "T", "C", "h", "w", "H", "W", "block", "k", "batch_size")
    m = dict(zip(keys, (int(v) for v in g["meta"])))
    gs = list(range(m["N"]))
    low = synth.pool_logits(m["seed"], gs, m["T"], m["C"], m["h"], m["w"], m["block"])
    labels = synth.pool_labels(m["seed"], gs, m["H"], m["W"], m["C"], m["block"] * 4)
    assert sha(low) == str(g["lowres_sha"]) and sha(labels) == str(g["labels_sha"]), \
        "synthetic input stream changed - regenerate goldens"
    return g, m, low, labels
