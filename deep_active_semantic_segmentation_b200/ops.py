"""Tensor-level wrappers over the C ABI (include/das_b200.h).

PyTorch is used for device memory and streams only: every function takes CUDA tensors, extracts
`data_ptr()` / the current stream and calls libdas_b200.so through ctypes.  Nothing here computes
scores on the host and nothing falls back to torch ops - a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, Sequence

import torch

from . import _lib
from ._lib import MC_PROBS, MC_SINGLE_SHOT, MC_VOTES, N_SCORES, DasError, McDesc, check

MAP_NAMES = ("vote_entropy", "pred_entropy", "bald", "confidence", "margin")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device=None):
    """torch's current stream ON THE TENSORS' DEVICE (not on whatever device happens to be current)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _h(device=None):
    """das_handle* of a device (one per device per process, _lib.handle)."""
    return _lib.handle(device)


import contextlib


@contextlib.contextmanager
def option(name: str, value: int, device=None):
    """Temporarily set a das_handle option, e.g. `with ops.option("mc_tma", 0): ...` (A/B measurements, tests)."""
    old = _lib.set_option(name, value, device)
    try:
        yield
    finally:
        _lib.set_option(name, old, device)


def _need_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DasError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise DasError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


class MCState:
    """Running Monte-Carlo state for one batch of B images (K1 accumulate + K2 finalize).

    Mirrors the life of `outputs = torch.cuda.FloatTensor(B, MC_STEPS, H, W)` in the reference
    (mc_dropout.py:37): created per batch, fed one stochastic forward at a time (or a group of them),
    finalised into entropy maps and per-image scores.
    """

    def __init__(self, B: int, C_: int, H: int, W: int, T_cap: int, votes: bool = True, probs: bool = True,
                 device=None, single_shot: bool = False):
        """single_shot: all T_cap (<= 32) passes will arrive in one score() call - the state then holds
        only the per-block partial sums (no accumulators, no votes)."""
        self.lib = _lib.load()
        self.single_shot = single_shot
        self.desc = McDesc(B, C_, H, W, T_cap, (MC_VOTES if votes else 0) | (MC_PROBS if probs else 0)
                           | (MC_SINGLE_SHOT if single_shot else 0))
        self.B, self.C, self.H, self.W, self.T_cap = B, C_, H, W, T_cap
        self.votes, self.probs = votes, probs
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.h = _h(self.device)
        nbytes = C.c_size_t()
        check(self.lib.das_mc_state_bytes(C.byref(self.desc), C.byref(nbytes)), "das_mc_state_bytes")
        self.nbytes = nbytes.value
        self.state = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.n_passes = 0

    def reset(self):
        self.n_passes = 0

    def accumulate(self, pass_logits: torch.Tensor | Sequence[torch.Tensor]):
        """Consume one pass ([B,C,H,W] logits) or a group of passes (sequence of such tensors)."""
        group = [pass_logits] if isinstance(pass_logits, torch.Tensor) else list(pass_logits)
        start = 0
        while start < len(group):
            chunk = self._check_group(group[start:start + _lib.MAX_PASS_GROUP])
            arr = (C.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
            check(self.lib.das_mc_accumulate(self.h, C.byref(self.desc), _ptr(self.state), arr, len(chunk), self.n_passes,
                                             _stream(self.device)), "das_mc_accumulate")
            self.n_passes += len(chunk)
            start += len(chunk)

    def _check_group(self, pass_logits):
        group = [pass_logits] if isinstance(pass_logits, torch.Tensor) else list(pass_logits)
        chunk = [_need_cuda(t, "logits", torch.float32) for t in group]
        for t in chunk:
            if t.device != self.device:
                raise DasError(f"logits live on {t.device}, the state on {self.device}")
            if tuple(t.shape) != (self.B, self.C, self.H, self.W):
                raise DasError(f"logits shape {tuple(t.shape)} != {(self.B, self.C, self.H, self.W)}")
        return chunk

    def score(self, pass_logits, labels: torch.Tensor | None, maps: Iterable[str] = (), scores: bool = True,
              weak_labels: bool = False, scores_out: torch.Tensor | None = None) -> dict:
        """Fused K1+K2 on the LAST group of passes (das_mc_accumulate_finalize): earlier groups, if any,
        went through accumulate(); with none (the whole MC stack in one call) no state touches HBM."""
        chunk = self._check_group(pass_logits)
        if len(chunk) > _lib.MAX_PASS_GROUP:     # leading passes through the streaming kernel
            head = len(chunk) - _lib.MAX_PASS_GROUP
            self.accumulate(chunk[:head])
            chunk = chunk[head:]
        labels, bufs, wl, sc = self._outputs(labels, maps, scores, weak_labels, scores_out)
        arr = (C.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
        check(self.lib.das_mc_accumulate_finalize(self.h, C.byref(self.desc), _ptr(self.state), arr, len(chunk),
                                                  self.n_passes, _ptr(labels), *[_ptr(bufs.get(n)) for n in MAP_NAMES],
                                                  _ptr(wl), _ptr(sc), _stream(self.device)), "das_mc_accumulate_finalize")
        self.n_passes += len(chunk)
        return self._pack(bufs, wl, sc)

    def score_upsampled(self, pass_lowres_logits, labels: torch.Tensor | None, maps: Iterable[str] = (),
                        scores: bool = True, weak_labels: bool = False,
                        scores_out: torch.Tensor | None = None) -> dict:
        """Fused K1+K2 on the model's LOW-RESOLUTION logits (das_mc_upsample_accumulate_finalize): every tensor
        is the `low_res_x` [B,C,h,w] of one stochastic forward, i.e. what models/deeplab.py:59 feeds to
        F.interpolate(..., size=(H,W), mode='bilinear', align_corners=True); the interpolation happens inside
        the kernel.  The whole Monte-Carlo stack of the batch (<= 32 passes) arrives in this one call."""
        if self.n_passes != 0:
            raise DasError("score_upsampled() takes the whole Monte-Carlo stack of a batch in one call")
        group = [pass_lowres_logits] if isinstance(pass_lowres_logits, torch.Tensor) else list(pass_lowres_logits)
        chunk = [_need_cuda(t, "low-res logits", torch.float32) for t in group]
        if not 1 <= len(chunk) <= min(_lib.MAX_PASS_GROUP, self.T_cap):
            raise DasError(f"score_upsampled: {len(chunk)} passes, need 1..{min(_lib.MAX_PASS_GROUP, self.T_cap)}")
        h, w = int(chunk[0].shape[-2]), int(chunk[0].shape[-1])
        for t in chunk:
            if tuple(t.shape) != (self.B, self.C, h, w):
                raise DasError(f"low-res logits shape {tuple(t.shape)} != {(self.B, self.C, h, w)}")
        labels, bufs, wl, sc = self._outputs(labels, maps, scores, weak_labels, scores_out)
        arr = (C.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
        check(self.lib.das_mc_upsample_accumulate_finalize(self.h, C.byref(self.desc), _ptr(self.state), arr, len(chunk),
                                                           h, w, _ptr(labels), *[_ptr(bufs.get(n)) for n in MAP_NAMES],
                                                           _ptr(wl), _ptr(sc), _stream(self.device)),
              "das_mc_upsample_accumulate_finalize")
        self.n_passes += len(chunk)
        return self._pack(bufs, wl, sc)

    def finalize(self, labels: torch.Tensor | None, maps: Iterable[str] = (), scores: bool = True,
                 weak_labels: bool = False, scores_out: torch.Tensor | None = None) -> dict:
        """-> {'scores': f32 [B,6] (column order _lib.SCORE_INDEX), <map name>: f32 [B,H,W], 'weak_labels': u8}.
        scores_out: optional contiguous f32 [B,6] destination (e.g. a slice of a pool-wide score table)."""
        if self.n_passes < 1:
            raise DasError("finalize() before any accumulate()")
        labels, bufs, wl, sc = self._outputs(labels, maps, scores, weak_labels, scores_out)
        check(self.lib.das_mc_finalize(self.h, C.byref(self.desc), _ptr(self.state), _ptr(labels), self.n_passes,
                                       *[_ptr(bufs.get(n)) for n in MAP_NAMES], _ptr(wl), _ptr(sc), _stream(self.device)),
              "das_mc_finalize")
        return self._pack(bufs, wl, sc)

    def _outputs(self, labels, maps, scores, weak_labels, scores_out):
        if labels is not None:
            labels = _need_cuda(labels, "labels", torch.float32)
            if tuple(labels.shape) != (self.B, self.H, self.W):
                raise DasError(f"labels shape {tuple(labels.shape)} != {(self.B, self.H, self.W)}")
        bufs = {}
        for name in maps:
            if name not in MAP_NAMES:
                raise DasError(f"unknown map {name!r}")
            bufs[name] = torch.empty((self.B, self.H, self.W), dtype=torch.float32, device=self.device)
        wl = torch.empty((self.B, self.H, self.W), dtype=torch.uint8, device=self.device) if weak_labels else None
        if scores_out is not None:
            if (not scores_out.is_cuda or scores_out.dtype != torch.float32 or not scores_out.is_contiguous()
                    or tuple(scores_out.shape) != (self.B, N_SCORES)):
                raise DasError("scores_out must be a contiguous CUDA float32 [B, 6] tensor")
            sc = scores_out
        else:
            sc = torch.empty((self.B, N_SCORES), dtype=torch.float32, device=self.device) if scores else None
        return labels, bufs, wl, sc

    @staticmethod
    def _pack(bufs, wl, sc):
        out = dict(bufs)
        if wl is not None:
            out["weak_labels"] = wl
        if sc is not None:
            out["scores"] = sc
        return out

    def votes_tensor(self) -> torch.Tensor:
        """u8 [B,T_cap,H,W] copy of the recorded votes (tests / debugging)."""
        p = C.c_void_p()
        check(self.lib.das_mc_votes_ptr(C.byref(self.desc), _ptr(self.state), C.byref(p)), "das_mc_votes_ptr")
        off = p.value - self.state.data_ptr()
        n = self.B * self.T_cap * self.H * self.W
        return self.state[off:off + n].view(self.B, self.T_cap, self.H, self.W).clone()


def upsample_supported(h: int, w: int, H: int, W: int, device=None) -> bool:
    """Can score_upsampled() interpolate h x w -> H x W in-kernel (else: F.interpolate + score())?"""
    return bool(_lib.load().das_mc_upsample_supported(_h(device), int(h), int(w), int(H), int(W)))


def upsample_variant(B: int, C: int, h: int, w: int, H: int, W: int, votes: bool = True, probs: bool = True,
                     device=None, default_options: bool = False) -> int:
    """Which fused-upsample kernel score_upsampled() would launch (das_mc_upsample_variant, host only): 0 = shape not
    supported, 4 / 15 = pixel pairs per lane, 220 / 216 = one pixel per lane (include/das_b200.h).
    `default_options=True` asks with a NULL handle (no CUDA context needed)."""
    flags = (MC_VOTES if votes else 0) | (MC_PROBS if probs else 0) | MC_SINGLE_SHOT
    desc = McDesc(int(B), int(C), int(H), int(W), 1, flags)
    hd = None if default_options else _h(device)
    import ctypes   # the module alias `C` is shadowed by the class count here
    return int(_lib.load().das_mc_upsample_variant(hd, ctypes.byref(desc), int(h), int(w)))


# ---------------------------------------------------------------------------------------------
# region path
# ---------------------------------------------------------------------------------------------

def suppress_rects(maps: torch.Tensor, rects) -> None:
    """In place: zero maps[i, r:r+h, c:c+w] for every (i, r, c, h, w) (mc_dropout.py:110-121)."""
    maps_c = _need_cuda(maps, "maps", torch.float32)
    if maps_c.data_ptr() != maps.data_ptr():
        raise DasError("maps must be contiguous (modified in place)")
    if len(rects) == 0:
        return
    B, H, W = maps.shape
    if isinstance(rects, torch.Tensor) and rects.is_cuda:
        r = rects.to(torch.int32).reshape(-1, 5).contiguous()
        check(_lib.load().das_suppress_rects(_h(maps.device), _ptr(maps), B, H, W, _ptr(r), r.shape[0],
                                             _stream(maps.device)), "das_suppress_rects")
        return
    # host records (the caller's lists): passed in the kernel parameters - no upload, no synchronisation
    flat = [int(v) for rec in (rects.tolist() if isinstance(rects, torch.Tensor) else rects) for v in rec]
    if len(flat) % 5:
        raise DasError("suppress_rects: records are (image, r, c, h, w)")
    arr = (C.c_int32 * len(flat))(*flat)
    check(_lib.load().das_suppress_rects_host(_h(maps.device), _ptr(maps), B, H, W, arr, len(flat) // 5,
                                              _stream(maps.device)), "das_suppress_rects_host")


def add_maps(a: torch.Tensor, b: torch.Tensor) -> None:
    _need_cuda(a, "a", torch.float32)
    b = _need_cuda(b, "b", torch.float32)
    if not a.is_contiguous() or a.numel() != b.numel():
        raise DasError("add_maps: a must be contiguous and the sizes must agree")
    check(_lib.load().das_add_maps(_h(a.device), _ptr(a), _ptr(b), a.numel(), _stream(a.device)), "das_add_maps")


def add_gaussian_noise(x: torch.Tensor, sigma: float, seed: int, stream_id: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """x + sigma * N(0,1) drawn on the device (das_add_gaussian_noise, mc_noise.py:26-27); the noise is a pure function
    of (seed, stream_id, element index).  `out` may be a reusable buffer of x's shape (or x itself)."""
    x = _need_cuda(x, "x", torch.float32)
    if out is None:
        out = torch.empty_like(x)
    elif not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.shape != x.shape:
        raise DasError("add_gaussian_noise: `out` must be a contiguous CUDA float32 tensor of x's shape")
    check(_lib.load().das_add_gaussian_noise(_h(x.device), _ptr(x), x.numel(), C.c_float(sigma), int(seed) & (2 ** 64 - 1),
                                             int(stream_id) & (2 ** 64 - 1), _ptr(out), _stream(x.device)),
          "das_add_gaussian_noise")
    return out


def new_minmax(device) -> torch.Tensor:
    mm = torch.empty(2, dtype=torch.float32, device=device)
    check(_lib.load().das_minmax_init(_h(mm.device), _ptr(mm), _stream(mm.device)), "das_minmax_init")
    return mm


def box_sum(maps: torch.Tensor, R: int, minmax: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Stride-1 valid RxR box sums of maps [B,H,W] -> [B,H-R+1,W-R+1]; folds min/max into `minmax`."""
    maps = _need_cuda(maps, "maps", torch.float32)
    B, H, W = maps.shape
    lib = _lib.load()
    nbytes = C.c_size_t()
    check(lib.das_box_sum_workspace_bytes(B, H, W, R, C.byref(nbytes)), "das_box_sum_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=maps.device)
    if out is None:
        out = torch.empty((B, H - R + 1, W - R + 1), dtype=torch.float32, device=maps.device)
    elif not out.is_contiguous() or tuple(out.shape) != (B, H - R + 1, W - R + 1):
        raise DasError("box_sum: bad `out`")
    check(lib.das_box_sum(_h(maps.device), _ptr(maps), B, H, W, R, _ptr(out), _ptr(minmax), _ptr(ws), _stream(maps.device)),
          "das_box_sum")
    return out


def minmax_normalise(score_maps: torch.Tensor, minmax: torch.Tensor) -> None:
    if not score_maps.is_contiguous():
        raise DasError("minmax_normalise: score_maps must be contiguous (modified in place)")
    _need_cuda(score_maps, "score_maps", torch.float32)
    check(_lib.load().das_minmax_normalise(_h(score_maps.device), _ptr(score_maps), score_maps.numel(), _ptr(minmax),
                                           _stream(score_maps.device)), "das_minmax_normalise")


def nms_pick_bound(H2: int, W2: int, R: int) -> int:
    """Picks of one image are pairwise >= R apart in the max-norm, so at most this many exist."""
    return math.ceil(H2 / R) * math.ceil(W2 / R)


def nms_sequences(score_maps: torch.Tensor, R: int, kmax: int, stop: float = 0.01, image_offset: int = 0,
                  with_flat: bool = False):
    """Per-image greedy NMS sequences; MUTATES score_maps. -> (scores [N,kmax], rc [N,kmax,2], count [N])
    (+ flat int64 [N,kmax] pool indices with with_flat=True).  Unused score slots are -inf."""
    if not score_maps.is_contiguous():
        raise DasError("nms_sequences: score_maps must be contiguous (modified in place)")
    _need_cuda(score_maps, "score_maps", torch.float32)
    N, H2, W2 = score_maps.shape
    dev = score_maps.device
    cs = torch.empty((N, kmax), dtype=torch.float32, device=dev)
    rc = torch.zeros((N, kmax, 2), dtype=torch.int32, device=dev)
    cnt = torch.zeros((N,), dtype=torch.int32, device=dev)
    flat = torch.empty((N, kmax), dtype=torch.int64, device=dev) if with_flat else None
    check(_lib.load().das_nms_sequences(_h(dev), _ptr(score_maps), N, H2, W2, R, kmax, C.c_float(stop), _ptr(cs), _ptr(rc),
                                        _ptr(cnt), int(image_offset), _ptr(flat), _stream(dev)), "das_nms_sequences")
    if with_flat:
        return cs, rc, cnt, flat
    return cs, rc, cnt


# ---------------------------------------------------------------------------------------------
# accuracy-predictor selectors
# ---------------------------------------------------------------------------------------------

def accuracy_scores(logits: torch.Tensor, labels: torch.Tensor | None, num_classes: int, p0_map: bool = False):
    """logits f32 [B,C,H,W], labels f32 [B,H,W] -> scores f32 [B,5] (column order _lib.ACC_INDEX)
    [+ softmax[0] map with invalid pixels zeroed, f32 [B,H,W]] (accuracy.py:30-33,55-64,117-118,159-162)."""
    logits = _need_cuda(logits, "logits", torch.float32)
    B, Cc, H, W = logits.shape
    if labels is not None:
        labels = _need_cuda(labels, "labels", torch.float32)
        if tuple(labels.shape) != (B, H, W):
            raise DasError(f"labels shape {tuple(labels.shape)} != {(B, H, W)}")
    lib = _lib.load()
    nbytes = C.c_size_t()
    check(lib.das_accuracy_workspace_bytes(B, H, W, C.byref(nbytes)), "das_accuracy_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=logits.device)
    scores = torch.empty((B, len(_lib.ACC_INDEX)), dtype=torch.float32, device=logits.device)
    pm = torch.empty((B, H, W), dtype=torch.float32, device=logits.device) if p0_map else None
    check(lib.das_accuracy_scores(_h(logits.device), _ptr(logits), B, Cc, H, W, _ptr(labels), int(num_classes), _ptr(pm),
                                  _ptr(scores), _ptr(ws), _stream(logits.device)), "das_accuracy_scores")
    return (scores, pm) if p0_map else scores


# ---------------------------------------------------------------------------------------------
# max-subset representativeness
# ---------------------------------------------------------------------------------------------

def maxsubset_greedy(X: torch.Tensor, Y: torch.Tensor, k: int) -> torch.Tensor:
    """Greedy facility location (max_subset.py:17-39): X [N,D] pool, Y [M,D] candidates (both float32 or both
    float64, CUDA) -> int32 [k] candidate indices in pick order."""
    if X.dtype != Y.dtype or X.dtype not in (torch.float32, torch.float64):
        raise DasError("maxsubset_greedy: X and Y must both be float32 or both float64")
    X, Y = _need_cuda(X, "X", X.dtype), _need_cuda(Y, "Y", Y.dtype)
    (N, D), (M, D2) = X.shape, Y.shape
    if D != D2:
        raise DasError("maxsubset_greedy: feature dimensions differ")
    f64 = 1 if X.dtype == torch.float64 else 0
    lib = _lib.load()
    nbytes = C.c_size_t()
    check(lib.das_maxsubset_workspace_bytes(N, M, D, f64, C.byref(nbytes)), "das_maxsubset_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=X.device)
    picks = torch.empty(max(k, 1), dtype=torch.int32, device=X.device)
    check(lib.das_maxsubset_greedy(_h(X.device), _ptr(X), _ptr(Y), N, M, D, f64, int(k), _ptr(picks), _ptr(ws),
                                   _stream(X.device)), "das_maxsubset_greedy")
    return picks[:k]


# ---------------------------------------------------------------------------------------------
# ranking
# ---------------------------------------------------------------------------------------------

def topk(scores: torch.Tensor, k: int, descending: bool, ids: torch.Tensor | None = None):
    """First k of the stable sort of `scores` -> (scores [k], ids int64 [k]); ids default to positions."""
    scores = _need_cuda(scores, "scores", torch.float32).reshape(-1)
    n = scores.numel()
    k = max(0, min(int(k), n))
    out_s = torch.empty(k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(k, dtype=torch.int64, device=scores.device)
    if k == 0:
        return out_s, out_i
    if ids is not None:
        ids = _need_cuda(ids, "ids", torch.int64).reshape(-1)
        if ids.numel() != n:
            raise DasError("topk: ids and scores differ in length")
    check(_lib.load().das_topk(_h(scores.device), _ptr(scores), _ptr(ids), n, k, 1 if descending else 0, _ptr(out_s),
                               _ptr(out_i), None, _stream(scores.device)), "das_topk")
    return out_s, out_i


def topk_records(scores: torch.Tensor, k_slots: int, descending: bool, ids: torch.Tensor | None = None,
                 id_offset: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """This rank's best min(k_slots, n) candidates as int64 records [k_slots, 2] = {float32 score bits, id + id_offset}
    (padding records {-/+inf, -1} behind them): the send buffer of the candidate all-gather (das_topk_records)."""
    scores = _need_cuda(scores, "scores", torch.float32).reshape(-1)
    n = scores.numel()
    if ids is not None:
        ids = _need_cuda(ids, "ids", torch.int64).reshape(-1)
        if ids.numel() != n:
            raise DasError("topk_records: ids and scores differ in length")
    rec = torch.empty((int(k_slots), 2), dtype=torch.int64, device=scores.device) if out is None else out
    check(_lib.load().das_topk_records(_h(scores.device), _ptr(scores) if n else None, _ptr(ids), n, int(k_slots),
                                       1 if descending else 0, int(id_offset), _ptr(rec), _stream(scores.device)),
          "das_topk_records")
    return rec


def topk_merge(records: torch.Tensor, k_slots: int, descending: bool) -> torch.Tensor:
    """First k_slots records of the stable ranking of a gathered record table [m, 2] (das_topk_merge)."""
    records = _need_cuda(records, "records", torch.int64)
    out = torch.empty((int(k_slots), 2), dtype=torch.int64, device=records.device)
    check(_lib.load().das_topk_merge(_h(records.device), _ptr(records), records.shape[0], int(k_slots),
                                     1 if descending else 0, _ptr(out), _stream(records.device)), "das_topk_merge")
    return out


def records_to_host(records: torch.Tensor):
    """int64 records [m, 2] (any device) -> (float32 scores [m'], int64 ids [m']) numpy arrays without the padding."""
    import numpy as np

    r = records.detach().cpu().numpy()
    keep = r[:, 1] >= 0
    sc = np.ascontiguousarray(r[keep, 0]).astype(np.uint32).view(np.float32)
    return sc, np.ascontiguousarray(r[keep, 1])


# ---------------------------------------------------------------------------------------------
# core-set
# ---------------------------------------------------------------------------------------------

class KCenterFilter:
    """Tensor-core distance filter for K4 (das_kcenter_filter_build): bf16 tcgen05 Gram distances of the row shard
    [row_begin,row_end) against every candidate centre.  Purely an accelerator - results are bit-identical
    with and without it."""

    def __init__(self, feats: torch.Tensor, row_begin: int = 0, row_end: int | None = None):
        feats = _need_cuda(feats, "feats", torch.float32)
        self.N, self.D = feats.shape
        self.row_begin, self.row_end = int(row_begin), self.N if row_end is None else int(row_end)
        self.rows = self.row_end - self.row_begin
        lib = _lib.load()
        nbytes = C.c_size_t()
        check(lib.das_kcenter_filter_bytes(self.N, self.D, self.rows, C.byref(nbytes)), "das_kcenter_filter_bytes")
        self.nbytes = nbytes.value
        raw = torch.empty(self.nbytes + 1024, dtype=torch.uint8, device=feats.device)
        off = (-raw.data_ptr()) % 1024
        self.blob = raw[off:off + self.nbytes]
        self.device = feats.device
        check(lib.das_kcenter_filter_build(_h(feats.device), _ptr(feats), self.N, self.D, self.row_begin, self.row_end,
                                           _ptr(self.blob), _stream(feats.device)), "das_kcenter_filter_build")

    @staticmethod
    def bytes_needed(N: int, D: int, rows: int) -> int:
        nbytes = C.c_size_t()
        check(_lib.load().das_kcenter_filter_bytes(N, D, rows, C.byref(nbytes)), "das_kcenter_filter_bytes")
        return nbytes.value

    def stats(self):
        """(exact float64 row evaluations, rows screened) so far."""
        out = (C.c_uint64 * 2)()
        check(_lib.load().das_kcenter_filter_stats(_h(self.device), _ptr(self.blob), self.N, self.D, self.rows, out,
                                                   _stream(self.device)), "das_kcenter_filter_stats")
        return int(out[0]), int(out[1])


def _filter_ptr(flt, feats, row_begin, row_end):
    if flt is None:
        return None
    N, D = feats.shape
    if (flt.N, flt.D, flt.row_begin, flt.row_end) != (N, D, row_begin, row_end):
        raise DasError("KCenterFilter was built for a different feature matrix / row shard")
    return _ptr(flt.blob)


def kcenter_greedy(feats: torch.Tensor, centers: Sequence[int] | torch.Tensor, K: int, flt: KCenterFilter | None = None):
    """Single-GPU k-center greedy -> (picks int32 [K], min_dist f64 [N]) on the device."""
    feats = _need_cuda(feats, "feats", torch.float32)
    N, D = feats.shape
    cen = torch.as_tensor(centers, dtype=torch.int32).reshape(-1).to(feats.device)
    lib = _lib.load()
    nbytes = C.c_size_t()
    check(lib.das_kcenter_workspace_bytes(_h(feats.device), N, D, C.byref(nbytes)), "das_kcenter_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=feats.device)
    picks = torch.empty(max(K, 1), dtype=torch.int32, device=feats.device)
    min_d = torch.empty(N, dtype=torch.float64, device=feats.device)
    check(lib.das_kcenter_greedy(_h(feats.device), _ptr(feats), N, D, _ptr(cen), cen.numel(), K, _ptr(picks), _ptr(min_d),
                                 _ptr(ws), _filter_ptr(flt, feats, 0, N), _stream(feats.device)), "das_kcenter_greedy")
    return picks[:K], min_d


def kcenter_init(feats, row_begin, row_end, centers, min_d2, key2, flt: KCenterFilter | None = None):
    N, D = feats.shape
    check(_lib.load().das_kcenter_init(_h(feats.device), _ptr(feats), N, D, row_begin, row_end, _ptr(centers),
                                       centers.numel(), _ptr(min_d2), _ptr(key2),
                                       _filter_ptr(flt, feats, row_begin, row_end), _stream(feats.device)),
          "das_kcenter_init")


def kcenter_step(feats, row_begin, row_end, centre_idx, min_d2, key2, flt: KCenterFilter | None = None):
    N, D = feats.shape
    check(_lib.load().das_kcenter_step(_h(feats.device), _ptr(feats), N, D, row_begin, row_end, _ptr(centre_idx),
                                       _ptr(min_d2), _ptr(key2), _filter_ptr(flt, feats, row_begin, row_end),
                                       _stream(feats.device)), "das_kcenter_step")


def kcenter_filter_budget_ok(N: int, D: int, rows: int, device) -> bool:
    """Use the tensor-core filter when its N x rows float32 table fits comfortably in free HBM."""
    free, _total = torch.cuda.mem_get_info(device)
    return KCenterFilter.bytes_needed(N, D, rows) <= free // 3
