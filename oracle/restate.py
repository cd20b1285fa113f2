"""TEST INFRASTRUCTURE - CPU restatement (numpy, float32) of the reference's selection-scoring
arithmetic.  It is the *checker* for the CUDA path, never the product: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_golden.py)
against tests/golden/*.npz, which were produced by running the reference's own, unmodified
selector classes in this container (oracle/gen_golden.py via oracle/ref_shim.py), and
against the reference's three data-free fixtures.
The composed scores that the reference never computes (predictive entropy of the MC mean,
BALD, MC confidence/margin - SURVEY.md F2/F3) are composed from the reference's own
primitives and are labelled "composed" below; their T=1 special cases are pinned by CEAL.

All file:line citations are relative to the reference tree.
"""
from __future__ import annotations

import math

import numpy as np

EPS = np.float32(1e-12)  # mc_dropout.py:48, ceal.py:118


# --------------------------------------------------------------------------------------
# per-pixel reductions
# --------------------------------------------------------------------------------------

def valid_mask(labels: np.ndarray, C: int) -> np.ndarray:
    """~((label < 0) | (label >= C)); labels are float32 (mc_dropout.py:45, ceal.py:40)."""
    return ~((labels < 0) | (labels >= C))


def votes_from_logits(logits: np.ndarray) -> np.ndarray:
    """Per-pass class vote = argmax over the class axis, first maximal index on ties
    (torch.argmax(model(x), dim=1), mc_dropout.py:40).  logits [..., C, H, W] -> uint8 [..., H, W]."""
    return np.argmax(logits, axis=-3).astype(np.uint8)


def vote_entropy_map(votes: np.ndarray, C: int, valid: np.ndarray | None) -> np.ndarray:
    """votes uint8 [T,H,W] -> float32 [H,W].

    p_c = n_c / T in float32; VE = sum over c ascending of -(p_c * log2(p_c + 1e-12)),
    accumulated in float32; invalid pixels -> 0 (mc_dropout.py:43-49; same loop in
    mc_noise.py:31-37, 72-78, 100-106)."""
    T = votes.shape[0]
    ve = np.zeros(votes.shape[1:], dtype=np.float32)
    for c in range(C):
        p = (votes == c).sum(axis=0).astype(np.float32) / np.float32(T)
        ve = ve - (p * np.log2(p + EPS)).astype(np.float32)
    if valid is not None:
        ve[~valid] = 0
    return ve


def softmax_classes(logits: np.ndarray) -> np.ndarray:
    """nn.Softmax2d over the class axis (ceal.py:34-36, 81-82, 111-112): max-subtracted, float32."""
    m = logits.max(axis=-3, keepdims=True)
    e = np.exp((logits - m).astype(np.float32)).astype(np.float32)
    return (e / e.sum(axis=-3, keepdims=True, dtype=np.float32)).astype(np.float32)


def entropy_map(p: np.ndarray, valid: np.ndarray | None) -> np.ndarray:
    """p float32 [C,H,W] -> E = sum_c ascending of -(p_c*log2(p_c+1e-12)); invalid -> 0 (ceal.py:116-119)."""
    e = np.zeros(p.shape[1:], dtype=np.float32)
    for c in range(p.shape[0]):
        e = e - (p[c] * np.log2(p[c] + EPS)).astype(np.float32)
    if valid is not None:
        e[~valid] = 0
    return e


def confidence_map(p: np.ndarray, valid: np.ndarray | None) -> np.ndarray:
    """max_c p; invalid -> 1 (ceal.py:36-39)."""
    m = p.max(axis=0).astype(np.float32)
    if valid is not None:
        m[~valid] = 1
    return m


def margin_map(p: np.ndarray, valid: np.ndarray | None) -> np.ndarray:
    """largest minus second largest class probability; invalid -> 1 (ceal.py:84-91)."""
    part = np.partition(p, p.shape[0] - 2, axis=0)
    m = (part[-1] - part[-2]).astype(np.float32)
    if valid is not None:
        m[~valid] = 1
    return m


def _fma32(a, b, c):
    """float32 fused multiply-add: a*b is exact in float64 (24 + 24 significand bits), one rounding to float64
    and one to float32 follow (a double rounding that differs from a true fma in < 1e-9 of the cases)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _align_corners_axis(n_in: int, n_out: int):
    """ATen's source index / weights of one axis for mode='bilinear', align_corners=True
    (aten/src/ATen/native/UpSample.h: area_pixel_compute_scale, area_pixel_compute_source_index,
    guard_index_and_lambda; PyTorch 2.11 - not vendored in the reference, called from models/deeplab.py:59):
    scale = float(in-1)/float(out-1) (0 if out == 1), src = scale*dst, i0 = min(int(src), in-1),
    l1 = clamp(src - i0, 0, 1), l0 = 1 - l1, i1 = i0 + (i0 < in-1)."""
    scale = np.float32(0) if n_out <= 1 else np.float32(n_in - 1) / np.float32(n_out - 1)
    src = (scale * np.arange(n_out, dtype=np.float32)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    l1 = np.clip((src - i0.astype(np.float32)).astype(np.float32), np.float32(0), np.float32(1))
    l0 = (np.float32(1) - l1).astype(np.float32)
    i1 = i0 + (i0 < n_in - 1)
    return i0, i1, l0, l1


def bilinear_upsample_align_corners(low: np.ndarray, H: int, W: int) -> np.ndarray:
    """F.interpolate(low, size=(H, W), mode='bilinear', align_corners=True) - the last op of the reference
    models (models/deeplab.py:59, unet.py:58, fastscnn.py:22).  low float32 [..., h, w] -> [..., H, W].

    out = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d), each sum evaluated as fma(l0, ., l1*.) - which is
    bit-identical to ATen's vectorised CPU kernel (torch 2.11, AVX-512) for outputs such as 17->65, 33->129,
    129->513 and 128x256->512x1024 (pinned by tests/golden/upsample_*.npz); on the tiny even-sized case in the
    goldens ATen takes a differently rounded code path that is within 2 float32 ulp of this one."""
    low = np.asarray(low, np.float32)
    h, w = low.shape[-2:]
    y0, y1, ly0, ly1 = _align_corners_axis(h, H)
    x0, x1, lx0, lx1 = _align_corners_axis(w, W)
    top, bot = low[..., y0, :], low[..., y1, :]
    t = _fma32(lx0, top[..., x0], (lx1 * top[..., x1]).astype(np.float32))
    b = _fma32(lx0, bot[..., x0], (lx1 * bot[..., x1]).astype(np.float32))
    return _fma32(ly0[:, None], t, (ly1[:, None] * b).astype(np.float32))


def mc_maps_upsampled(pass_lowres: np.ndarray, labels: np.ndarray | None, C: int, H: int, W: int) -> dict:
    """mc_maps() of the model output the reference selectors see: the low-resolution decoder logits
    [T,C,h,w] interpolated to H x W (models/deeplab.py:58-59), then mc_dropout.py:37-49 / ceal.py."""
    return mc_maps(bilinear_upsample_align_corners(pass_lowres, H, W), labels, C)


def mc_maps(pass_logits: np.ndarray, labels: np.ndarray | None, C: int) -> dict:
    """All per-pixel maps for one image.  pass_logits float32 [T,C,H,W].

    vote_entropy           - reference, mc_dropout.py:37-49
    entropy/conf/margin    - reference when T == 1 (ceal.py); for T > 1 they are evaluated
                             on the MC mean  p_bar = (1/T) sum_t softmax(x_t)   [composed]
    expected_entropy, bald - [composed]  BALD = E(p_bar) - (1/T) sum_t E(p_t)
    """
    T = pass_logits.shape[0]
    valid = valid_mask(labels, C) if labels is not None else None
    votes = votes_from_logits(pass_logits)
    p_t = softmax_classes(pass_logits)                       # [T,C,H,W]
    acc = np.zeros(p_t.shape[1:], dtype=np.float32)
    eacc = np.zeros(p_t.shape[2:], dtype=np.float32)
    for t in range(T):                                       # running float32 accumulators
        acc = acc + p_t[t]
        eacc = eacc + entropy_map(p_t[t], None)
    p_bar = (acc / np.float32(T)).astype(np.float32)
    pe = entropy_map(p_bar, valid)
    ee = (eacc / np.float32(T)).astype(np.float32)
    if valid is not None:
        ee[~valid] = 0
    return {
        "votes": votes,
        "vote_entropy": vote_entropy_map(votes, C, valid),
        "pred_entropy": pe,
        "expected_entropy": ee,
        "bald": (pe - ee).astype(np.float32),
        "confidence": confidence_map(p_bar, valid),
        "margin": margin_map(p_bar, valid),
    }


SCORE_NAMES = ("vote_entropy", "pred_entropy", "bald", "confidence", "margin", "expected_entropy")


def image_scores(maps: dict) -> dict:
    """Image score = mean over ALL H*W pixels, masked values included
    (mc_dropout.py:189 torch.mean; mc_noise.py:56 sum/(H*W); ceal.py:59,95,123)."""
    return {k: np.float32(np.mean(maps[k], dtype=np.float64)) for k in SCORE_NAMES}


# --------------------------------------------------------------------------------------
# ranking
# --------------------------------------------------------------------------------------

def rank_topk(scores, k: int, descending: bool) -> list:
    """Indices of the first k entries of a *stable* sort on score - Python's sorted() on
    (score, item) pairs keyed by score; `reverse=True` keeps equal scores in input order
    (mc_dropout.py:195, ceal.py:69,97,130, mc_noise.py:59,128,147)."""
    order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=descending)
    return order[:k]


# --------------------------------------------------------------------------------------
# region path
# --------------------------------------------------------------------------------------

def suppress_rects(m: np.ndarray, rects) -> np.ndarray:
    """Zero [r:r+h, c:c+w] for each already-labelled (r,c,h,w), in place (mc_dropout.py:110-121)."""
    if rects:
        for (r, c, h, w) in rects:
            m[r:r + h, c:c + w] = 0
    return m


def box_sum(m: np.ndarray, R: int) -> np.ndarray:
    """Stride-1 'valid' RxR box sum = conv2d with a ones kernel (mc_dropout.py:148-149).

    Evaluated exactly (float64 summed-area table) and rounded once to float32; the
    reference's float32 convolution agrees to its own rounding (checked in the golden test)."""
    H, W = m.shape
    sat = np.zeros((H + 1, W + 1), dtype=np.float64)
    sat[1:, 1:] = np.cumsum(np.cumsum(m.astype(np.float64), axis=0), axis=1)
    s = sat[R:, R:] - sat[:-R, R:] - sat[R:, :-R] + sat[:-R, :-R]
    return s.astype(np.float32)


def minmax_normalise(score_maps: np.ndarray, mn=None, mx=None) -> np.ndarray:
    """x.add_(-min).mul_(1/(max-min)) with a float32 scalar reciprocal (mc_dropout.py:152-155)."""
    mn = np.float32(score_maps.min() if mn is None else mn)
    mx = np.float32(score_maps.max() if mx is None else mx)
    inv = np.float32(1.0) / np.float32(mx - mn)
    return ((score_maps + np.float32(-mn)).astype(np.float32) * inv).astype(np.float32)


NMS_STOP = np.float32(0.01)  # mc_dropout.py:105 (float32 tensor compared with the scalar 0.01)


def square_nms(score_maps: np.ndarray, R: int, max_selection_count: float):
    """Sequential global greedy NMS, as written (mc_dropout.py:82-108); mutates score_maps.

    <= ceil(K) iterations: first flat argmax over the whole pool -> (i,r,c); record
    (r,c,R,R); zero [r-R,r+R) x [c-R,c+R) of image i; stop once the pool max < 0.01."""
    N, H2, W2 = score_maps.shape
    selected = [[] for _ in range(N)]
    count = 0
    flat = score_maps.reshape(-1)
    for _ in range(math.ceil(max_selection_count)):
        a = int(np.argmax(flat))
        i, r, c = a // (H2 * W2), (a // W2) % H2, a % W2
        selected[i].append((r, c, R, R))
        count += 1
        score_maps[i, max(0, r - R):min(H2, r + R), max(0, c - R):min(W2, c + R)] = 0
        if flat.max() < NMS_STOP:
            break
    return selected, count


def nms_sequence_single(score_map: np.ndarray, R: int, kmax: int):
    """Greedy NMS picks of ONE image, in pick order: list of (score, r, c).  Mutates the map.

    The first pick is always taken; later picks only while the map max is >= 0.01 - these
    are exactly the picks the global loop could ever take from this image."""
    H2, W2 = score_map.shape
    out = []
    flat = score_map.reshape(-1)
    while len(out) < kmax:
        a = int(np.argmax(flat))
        s = flat[a]
        if out and s < NMS_STOP:
            break
        r, c = a // W2, a % W2
        out.append((np.float32(s), r, c))
        score_map[max(0, r - R):min(H2, r + R), max(0, c - R):min(W2, c + R)] = 0
    return out


def merge_nms_sequences(seqs, R: int, max_selection_count: float, H2: int, W2: int):
    """k-way merge of per-image pick sequences == square_nms.

    Global order: score descending, then flat pool index (image, r, c) ascending - the
    first-flat-argmax rule of mc_dropout.py:91.  Within an image the sequence order is
    already forced (a later pick only exists after the earlier one suppressed its window).
    Stop rule: pick j+1 is taken iff j < ceil(K) and its score (= pool max after pick j) >= 0.01."""
    import heapq

    heap = []
    for i, seq in enumerate(seqs):
        if seq:
            s, r, c = seq[0]
            heapq.heappush(heap, (-float(s), (i * H2 + r) * W2 + c, i, 0))
    selected = [[] for _ in seqs]
    count = 0
    kmax = math.ceil(max_selection_count)
    while heap and count < kmax:
        negs, _, i, j = heapq.heappop(heap)
        if count > 0 and np.float32(-negs) < NMS_STOP:
            break
        s, r, c = seqs[i][j]
        selected[i].append((r, c, R, R))
        count += 1
        if j + 1 < len(seqs[i]):
            s2, r2, c2 = seqs[i][j + 1]
            heapq.heappush(heap, (-float(s2), (i * H2 + r2) * W2 + c2, i, j + 1))
    return selected, count


def region_selection(ve_maps, existing_regions, R: int, selection_size: int, base_size: int):
    """create_region_maps after the per-batch scoring (mc_dropout.py:144-158): suppress ->
    box-sum -> pool-global min-max -> K = selection_size*base^2/R^2 -> NMS."""
    maps = np.stack([box_sum(suppress_rects(m.copy(), existing_regions[i]), R)
                     for i, m in enumerate(ve_maps)])
    norm = minmax_normalise(maps)
    K = (selection_size * base_size * base_size) / (R * R)
    return square_nms(norm.copy(), R, K), norm


# --------------------------------------------------------------------------------------
# core-set
# --------------------------------------------------------------------------------------

def _euclidean_f64(x: np.ndarray, xx: np.ndarray, y: np.ndarray) -> np.ndarray:
    """The formula of euclidean_to on float64 rows x with their squared norms xx already at hand."""
    d2 = xx[:, None] + (y * y).sum(1)[None, :] - 2.0 * (x @ y.T)
    np.maximum(d2, 0, out=d2)
    return np.sqrt(d2)


def euclidean_to(features: np.ndarray, centres: np.ndarray) -> np.ndarray:
    """sklearn pairwise_distances(metric='euclidean') on float64 input (core_set.py:34):
    sqrt(max(|x|^2 + |y|^2 - 2 x.y, 0)).  scikit-learn is a third-party dependency of the
    reference (unpinned there; 1.9.0 in this container) - this is its published formula."""
    x = np.asarray(features, dtype=np.float64)
    y = np.asarray(centres, dtype=np.float64)
    return _euclidean_f64(x, (x * x).sum(1), y)


def kcenter_greedy(features: np.ndarray, selected_indices, K: int):
    """_select_batch (core_set.py:17-30): m = min over already-selected of d; K times
    {j = first argmax m; assert j not selected; m = min(m, d(., j))}.  Returns (picks, m).
    The float64 copy of the rows and their squared norms are the same values at every step, so they are
    formed once (the per-step arithmetic is euclidean_to's, operand for operand)."""
    sel = list(selected_indices)
    x = np.asarray(features, dtype=np.float64)
    xx = (x * x).sum(1)
    m = _euclidean_f64(x, xx, x[sel]).min(axis=1)
    picks = []
    for _ in range(K):
        j = int(np.argmax(m))
        assert j not in sel, "k-center picked an already selected index (core_set.py:25)"
        m = np.minimum(m, _euclidean_f64(x, xx, x[[j]])[:, 0])
        picks.append(j)
    return picks, m


def avg_pool_features(feat: np.ndarray, k: int) -> np.ndarray:
    """F.avg_pool2d(feat, (k,k), k//2) flattened channel-major (core_set.py:56-63). feat [F,h,w]."""
    s = k // 2
    F_, h, w = feat.shape
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    out = np.empty((F_, oh, ow), dtype=np.float32)
    for i in range(oh):
        for j in range(ow):
            out[:, i, j] = feat[:, i * s:i * s + k, j * s:j * s + k].mean(axis=(1, 2), dtype=np.float64)
    return out.reshape(-1)


# --------------------------------------------------------------------------------------
# accuracy-predictor selectors (active_selection/accuracy.py)
# --------------------------------------------------------------------------------------

def accuracy_scores(seg_logits, unet_logits, labels: np.ndarray, num_classes: int) -> dict:
    """Per-image scores of ActiveSelectionAccuracy for one image.
    seg_logits [C,H,W] or None: wrong_count = #valid pixels with label != argmax (accuracy.py:30-33).
    unet_logits [2,H,W] or None: p0_sum = sum_valid softmax[0] (:55-58); not_argmax_sum = sum_valid (1 - argmax)
    (:60-64); unsure_mean = mean_valid (4 p1 - 4 p1^2) (:117-118).  valid = 0 <= label < num_classes."""
    valid = (labels >= 0) & (labels < num_classes)
    out = {"valid_count": np.float32(valid.sum())}
    if seg_logits is not None:
        pred = np.argmax(seg_logits, axis=0).astype(np.float32)
        out["wrong_count"] = np.float32((labels[valid] != pred[valid]).sum())
    if unet_logits is not None:
        p = softmax_classes(unet_logits)
        out["p0_sum"] = np.float32(p[0][valid].astype(np.float32).sum(dtype=np.float32))
        out["not_argmax_sum"] = np.float32((np.float32(1) - np.argmax(unet_logits, axis=0).astype(np.float32))[valid].sum(dtype=np.float32))
        p1 = p[1][valid].astype(np.float32)
        y = np.float32(4) * p1 - np.float32(4) * p1 ** 2
        out["unsure_mean"] = np.float32(y.mean(dtype=np.float32)) if y.size else np.float32(np.nan)
    return out


def accuracy_error_map(unet_logits: np.ndarray, labels: np.ndarray, num_classes: int) -> np.ndarray:
    """softmax[0] of the error head with invalid pixels zeroed (accuracy.py:159-162), float32 [H,W]."""
    p0 = softmax_classes(unet_logits)[0].astype(np.float32).copy()
    p0[(labels < 0) | (labels >= num_classes)] = 0
    return p0


# --------------------------------------------------------------------------------------
# max-subset representativeness (active_selection/max_subset.py)
# --------------------------------------------------------------------------------------

def max_representative_samples(image_features, candidate_features, selection_count: int) -> list:
    """Greedy facility location of max_subset.py:17-39.  D = sklearn pairwise_distances(X, Y) (euclidean; float32
    inputs give float32 distances, float64 inputs float64); repeat selection_count times: for every not yet
    selected candidate i, score_i = -sum_n min(min_d[n], D[n,i]); take the FIRST i with the strictly largest score;
    min_d = min(min_d, D[:, i])."""
    from sklearn.metrics import pairwise_distances

    X, Y = np.asarray(image_features), np.asarray(candidate_features)
    D = pairwise_distances(X, Y, metric="euclidean")
    min_d = np.ones(len(X)) * float("inf")
    selected = []
    for _ in range(selection_count):
        best, best_i, best_d = float("-inf"), None, None
        for i in range(len(Y)):
            if i in selected:
                continue
            tmp = np.minimum(min_d, D[:, i])
            score = np.sum(tmp) * -1
            if score > best:
                best, best_i, best_d = score, i, tmp
        selected.append(best_i)
        min_d = best_d
    return selected
