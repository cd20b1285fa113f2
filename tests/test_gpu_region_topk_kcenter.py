"""GPU parity of the region kernels (suppress / box sum / min-max / NMS), K3 top-k and K4 k-center."""
import numpy as np
import pytest
import torch

from deep_active_semantic_segmentation_b200 import synth
from oracle import restate as R
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _ops():
    from deep_active_semantic_segmentation_b200 import ops
    return ops


# ------------------------------------------------------------------ top-k

def _check_topk(scores, k, descending):
    ops = _ops()
    s, i = ops.topk(torch.tensor(scores, dtype=torch.float32).cuda(), k, descending)
    want = R.rank_topk([float(np.float32(v)) for v in scores], k, descending)
    assert i.cpu().tolist() == want                                   # index work: bit exact
    np.testing.assert_array_equal(s.cpu().numpy(), np.asarray(scores, dtype=np.float32)[want])


@pytest.mark.parametrize("descending", [True, False])
def test_topk_ties_negatives_and_edges(descending):
    _check_topk([0.5, 0.25, 0.5, 0.75, 0.25, 0.5], 4, descending)
    _check_topk([0.5, 0.25, 0.5, 0.75, 0.25, 0.5], 100, descending)       # k > n clamps
    _check_topk([1.0], 1, descending)
    _check_topk([-1.0, 0.0, -0.0, 3.0, -2.5, 0.0, -1.0], 7, descending)    # -0.0 == 0.0 is a tie
    _check_topk([0.0] * 300, 17, descending)                               # all equal: pure index order
    rng = np.random.default_rng(0)
    _check_topk(np.round(rng.random(2975), 2).tolist(), 125, descending)   # pool size of config 2, many ties
    _check_topk(rng.standard_normal(100_000).astype(np.float32).tolist(), 4096, descending)
    ops = _ops()
    s, i = ops.topk(torch.zeros(0, device="cuda"), 5, descending)
    assert s.numel() == 0 and i.numel() == 0


def test_topk_payload_ids_and_limits():
    ops = _ops()
    from deep_active_semantic_segmentation_b200._lib import DasError
    sc = torch.tensor([3.0, 1.0, 2.0, 3.0], device="cuda")
    ids = torch.tensor([40, 10, 20, 30], dtype=torch.int64, device="cuda")
    s, i = ops.topk(sc, 3, True, ids)
    assert i.cpu().tolist() == [40, 30, 20]        # stable on POSITION, payload carried along
    with pytest.raises(DasError):
        ops.topk(torch.zeros(10000, device="cuda"), 5000, True)     # > DAS_TOPK_MAX_K
    with pytest.raises(DasError):
        ops.topk(torch.zeros(4), 2, True)                           # CPU tensor


# ------------------------------------------------------------------ region path

def test_suppress_and_box_sum_match_oracle():
    ops = _ops()
    rng = np.random.default_rng(1)
    for (B, H, W, Rg) in [(3, 65, 65, 17), (2, 40, 100, 8), (1, 33, 47, 33), (2, 16, 16, 1), (1, 130, 1100, 64)]:
        maps = (rng.random((B, H, W)) * 4).astype(np.float32)
        maps[rng.random(maps.shape) < 0.4] = 0
        rects = [(0, 3, 5, 7, 9), (B - 1, H - 4, W - 6, 10, 10), (0, 0, 0, 2, W)]   # second one is clipped
        d = torch.from_numpy(maps).cuda()
        ops.suppress_rects(d, rects)
        want_m = maps.copy()
        for (i, r, c, h, w) in rects:
            R.suppress_rects(want_m[i], [(r, c, h, w)])
        np.testing.assert_array_equal(d.cpu().numpy(), want_m)
        mm = ops.new_minmax("cuda")
        out = ops.box_sum(d, Rg, mm).cpu().numpy()
        want = np.stack([R.box_sum(want_m[b], Rg) for b in range(B)])
        # fp64 sliding sums rounded once: equal to the exact oracle up to one float32 ulp
        np.testing.assert_allclose(out, want, rtol=1.2e-7, atol=1e-9)
        np.testing.assert_array_equal(mm.cpu().numpy(), np.array([out.min(), out.max()], dtype=np.float32))
        # min/max accumulate over batches (mc_dropout.py:152-153 on the whole pool)
        ops.box_sum(d * 2, Rg, mm)
        np.testing.assert_array_equal(mm.cpu().numpy(), np.array([out.min(), 2 * out.max()], dtype=np.float32))
        norm = torch.from_numpy(out).cuda()
        mm2 = torch.tensor([out.min(), out.max()], device="cuda")
        ops.minmax_normalise(norm, mm2)
        np.testing.assert_array_equal(norm.cpu().numpy(), R.minmax_normalise(out))     # same float32 ops


def _nms_gpu(maps, Rg, K):
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCDropout
    t = torch.from_numpy(maps.copy())
    regions, count = ActiveSelectionMCDropout.square_nms(t, Rg, K)
    return regions, count, t.numpy()


def test_nms_matches_sequential_reference_loop():
    rng = np.random.default_rng(5)
    for trial in range(10):
        N, H2, W2, Rg = int(rng.integers(1, 6)), int(rng.integers(6, 40)), int(rng.integers(6, 40)), int(rng.integers(2, 9))
        m = rng.random((N, H2, W2)).astype(np.float32)
        m[rng.random(m.shape) < 0.3] = 0
        if trial % 3 == 0:
            m = np.round(m, 1)               # many exact ties: first flat index must win
        K = float(rng.integers(1, 40)) + 0.5
        want = m.copy()
        ref_regions, ref_count = R.square_nms(want, Rg, K)
        regions, count, mutated = _nms_gpu(m, Rg, K)
        assert (regions, count) == (ref_regions, ref_count)
        np.testing.assert_array_equal(mutated, want)          # the caller's tensor is mutated like the reference's
    # degenerate: everything below the stop threshold -> exactly one pick
    regions, count, _ = _nms_gpu(np.full((3, 5, 5), 0.001, np.float32), 2, 4.0)
    assert count == 1 and regions[0] == [(0, 0, 2, 2)]


def test_nms_png_fixture_on_gpu():
    """The reference's own data-free NMS case (active_selection/tests.py:213-231) through the CUDA path."""
    ops = _ops()
    g = G.load("nms_png")
    Rg, K = int(g["R"]), int(g["K"])
    d = torch.from_numpy(g["images"].astype(np.float32) / 256).cuda()
    mm = ops.new_minmax("cuda")
    maps = ops.box_sum(d, Rg, mm)
    np.testing.assert_array_equal(maps.cpu().numpy(), g["raw_maps"])       # k/256 sums are exact
    assert float(mm[1]) == 8860.890625
    ops.minmax_normalise(maps, mm)
    np.testing.assert_array_equal(maps.cpu().numpy(), g["norm_maps"])
    regions, count, _ = _nms_gpu(maps.cpu().numpy(), Rg, K)
    assert count == 10 and regions == G.regions_from_rows(g["regions"], 2)


def test_nms_full_size_properties():
    """Cityscapes-shaped extension (512x1024, R=128): picks of an image are pairwise >= R apart (max-norm),
    scores are non-increasing along each sequence, every pick is the map maximum at its time."""
    ops = _ops()
    rng = np.random.default_rng(2)
    m = rng.random((4, 385, 897)).astype(np.float32)
    d = torch.from_numpy(m).cuda()
    kmax = ops.nms_pick_bound(385, 897, 128)
    cs, rc, cnt = ops.nms_sequences(d, 128, kmax, 0.01)
    cs, rc, cnt = cs.cpu().numpy(), rc.cpu().numpy(), cnt.cpu().numpy()
    for i in range(4):
        n = cnt[i]
        assert 1 <= n <= kmax
        assert cs[i, 0] == m[i].max() and (np.diff(cs[i, :n]) <= 0).all()
        pts = rc[i, :n]
        dist = np.abs(pts[:, None, :] - pts[None, :, :]).max(-1) + np.eye(n, dtype=np.int64) * 10**6
        assert dist.min() >= 128
        seq = R.nms_sequence_single(m[i].copy(), 128, kmax)
        assert [(r, c) for _, r, c in seq] == [tuple(p) for p in pts.tolist()]


# ------------------------------------------------------------------ k-center

def test_kcenter_toy_fixture_gpu():
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionCoreSet
    g = G.load("kcenter_toy")
    sel = ActiveSelectionCoreSet(None, None, None)
    assert sel._select_batch(g["features"], [6], 5) == [0, 2, 8, 4, 7]
    assert abs(float(sel.last_min_distances.max()) - 1.41421) < 1e-5


@pytest.mark.parametrize("name", ["coreset_small", "coreset_mid"])
def test_kcenter_matches_reference_golden(name):
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionCoreSet
    g = G.load(name)
    seed, N, D, L, K = (int(v) for v in g["meta"])
    feats = synth.coreset_features(seed, N, D)
    assert G.sha(feats) == str(g["features_sha"])
    sel = ActiveSelectionCoreSet(None, None, None)
    picks = sel._select_batch(feats.astype(np.float64), list(range(L)), K)
    assert picks == g["picks"].tolist()                                   # index work: bit exact
    # golden distances carry sklearn's cancellation noise where the true distance is 0 (see oracle test)
    np.testing.assert_allclose(sel.last_min_distances.cpu().numpy(), g["min_dist"], rtol=1e-9, atol=1e-4)


def test_kcenter_asserts_like_reference_and_rounds_wide_features_with_a_warning():
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionCoreSet
    sel = ActiveSelectionCoreSet(None, None, None)
    with pytest.raises(AssertionError):
        sel._select_batch(np.zeros((4, 3)), [0], 1)          # all distances 0 -> argmax 0, already selected
    # genuinely float64 features (0.1 is not a float32 value): accepted like the reference does, rounded, announced
    rng = np.random.default_rng(5)
    f64 = rng.standard_normal((40, 7)) + 0.1
    with pytest.warns(RuntimeWarning, match="rounded to float32"):
        picks = sel._select_batch(f64, [0, 1], 5)
    assert picks == R.kcenter_greedy(f64.astype(np.float32), [0, 1], 5)[0]


def test_updated_distances_keeps_the_reference_contract():
    """core_set.py:32-38: [n,1] min over the centres when min_distances is None, else np.minimum(min_distances, dist)."""
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionCoreSet
    sel = ActiveSelectionCoreSet(None, None, None)
    feats = synth.coreset_features(3, 200, 33).astype(np.float64)
    d0 = sel._updated_distances([4, 9, 17], feats, None)
    want0 = R.euclidean_to(feats, feats[[4, 9, 17]]).min(axis=1).reshape(-1, 1)
    assert isinstance(d0, np.ndarray) and d0.shape == (200, 1) and d0.dtype == np.float64
    np.testing.assert_allclose(d0, want0, rtol=1e-9, atol=1e-5)   # sklearn's expanded form carries ~1e-6 of noise at d = 0
    d1 = sel._updated_distances([50], feats, d0)
    np.testing.assert_allclose(d1, np.minimum(want0, R.euclidean_to(feats, feats[[50]])), rtol=1e-9, atol=1e-5)
    assert d1.shape == (200, 1) and float(d1[50, 0]) == 0.0 and float(d0[4, 0]) == 0.0
    d2 = sel._updated_distances([50, 60], feats, d0)              # broadcast like numpy does
    assert d2.shape == (200, 2)
    with pytest.raises(ValueError):
        sel._updated_distances([], feats, None)


def test_kcenter_odd_dimension_and_baseline_size_properties():
    ops = _ops()
    rng = np.random.default_rng(3)
    f = rng.standard_normal((257, 37)).astype(np.float32)    # D % 4 != 0 -> scalar path
    picks, md = ops.kcenter_greedy(torch.from_numpy(f).cuda(), [5, 9], 20)
    want, wm = R.kcenter_greedy(f, [5, 9], 20)
    assert picks.cpu().tolist() == want
    np.testing.assert_allclose(md.cpu().numpy(), wm, rtol=1e-9, atol=1e-6)
    # BASELINE config 5 size: N=10000, D=2048, L=50, K=500
    N, D, L, K = 10000, 2048, 50, 500
    feats = torch.from_numpy(synth.coreset_features(11, N, D)).cuda()
    picks, md = ops.kcenter_greedy(feats, list(range(L)), K)
    p = picks.cpu().tolist()
    assert len(set(p)) == K and not (set(p) & set(range(L)))
    md = md.cpu().numpy()
    assert (md[p] == 0).all() and (md[:L] == 0).all()
    # greedy property: the radius after the last pick is <= the distance at which the last pick was taken
    sub = feats[p + list(range(L))].double()
    d_all = torch.cdist(feats.double(), sub).min(dim=1).values.cpu().numpy()
    np.testing.assert_allclose(md, d_all, rtol=1e-9, atol=1e-5)


# ------------------------------------------------------------------ k-center, tensor-core distance filter

def _filter_table(flt):
    """(dt [N, ld] float32 view, exact float64 squared norms [N]) of a KCenterFilter blob (layout: gram.cuh)."""
    Dp = -(-flt.D // 64) * 64
    ld = -(-flt.rows // 32) * 32
    a = lambda x, al: -(-x // al) * al
    off_n64 = a(flt.N * Dp * 2, 1024)
    off_n32 = off_n64 + a(flt.N * 8, 256)
    off_dt = off_n32 + a(flt.N * 4, 256) + 256
    dt = flt.blob[off_dt:off_dt + flt.N * ld * 4].view(torch.float32).view(flt.N, ld)
    n64 = flt.blob[off_n64:off_n64 + flt.N * 8].view(torch.float64)
    return dt, n64


@pytest.mark.parametrize("N,D,lo,hi", [(300, 64, 0, 300), (1000, 200, 0, 1000), (777, 130, 129, 650), (2048, 2048, 0, 2048)])
def test_gram_filter_distances_respect_the_proven_bound(N, D, lo, hi):
    """tcgen05 bf16 distances vs exact float64: |dt - d2| <= 2^-7 (|a|^2 + |b|^2) for every (centre, row) pair."""
    ops = _ops()
    f = synth.coreset_features(5, N, D)
    feats = torch.from_numpy(f).cuda()
    flt = ops.KCenterFilter(feats, lo, hi)
    dt, n64 = _filter_table(flt)
    fd = feats.double()
    exact = torch.cdist(fd, fd[lo:hi]) ** 2                      # [N centres, rows]
    nrm = (fd * fd).sum(1)
    np.testing.assert_allclose(n64.cpu().numpy(), nrm.cpu().numpy(), rtol=1e-12)
    bound = (nrm[:, None] + nrm[None, lo:hi]) / 128.0
    err = (dt[:, :hi - lo].double() - exact).abs()
    assert bool((err <= bound).all()), float((err / bound).max())
    # and the bound is not vacuous: bf16 errors are far below it but above float32 noise
    assert float((err / bound).max()) < 0.5


@pytest.mark.parametrize("N,D,L,K", [(300, 64, 3, 40), (1500, 257, 7, 100), (4096, 512, 50, 200)])
def test_kcenter_filtered_is_bit_identical_to_exact(N, D, L, K):
    ops = _ops()
    feats = torch.from_numpy(synth.coreset_features(7, N, D)).cuda()
    p0, m0 = ops.kcenter_greedy(feats, list(range(L)), K)
    flt = ops.KCenterFilter(feats)
    p1, m1 = ops.kcenter_greedy(feats, list(range(L)), K, flt)
    assert p0.cpu().tolist() == p1.cpu().tolist()
    assert torch.equal(m0, m1)                                   # float64 min-distances: bit identical
    exact, screened = flt.stats()
    assert screened == K * N and 0 < exact < screened + N * L     # the filter did skip rows
    want, wm = R.kcenter_greedy(feats.cpu().numpy(), list(range(L)), K)
    assert p1.cpu().tolist() == want


def test_kcenter_filtered_adversarial_duplicates_and_near_ties():
    """Duplicated rows, rows at distance ~1e-4 of each other and a constant cluster: the filter must flag them
    all and the float64 result must not change."""
    ops = _ops()
    rng = np.random.default_rng(9)
    base = rng.standard_normal((64, 96)).astype(np.float32)
    f = np.concatenate([base, base, base + np.float32(1e-4) * rng.standard_normal((64, 96)).astype(np.float32),
                        np.ones((40, 96), np.float32)]).astype(np.float32)
    feats = torch.from_numpy(f).cuda()
    p0, m0 = ops.kcenter_greedy(feats, [0, 1], 60)
    p1, m1 = ops.kcenter_greedy(feats, [0, 1], 60, ops.KCenterFilter(feats))
    assert p0.cpu().tolist() == p1.cpu().tolist() and torch.equal(m0, m1)


def test_kcenter_sharded_steps_with_filter_match_single_shot():
    """The step-wise entry points (multi-GPU path) on two row shards of one device, with per-shard filters."""
    ops = _ops()
    N, D, L, K = 1200, 160, 5, 50
    feats = torch.from_numpy(synth.coreset_features(13, N, D)).cuda()
    want, wm = ops.kcenter_greedy(feats, list(range(L)), K)
    shards = [(0, 700), (700, N)]
    flts = [ops.KCenterFilter(feats, lo, hi) for lo, hi in shards]
    cen = torch.arange(L, dtype=torch.int32, device="cuda")
    md = [torch.empty(hi - lo, dtype=torch.float64, device="cuda") for lo, hi in shards]
    keys = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in shards]
    for (lo, hi), m, k, fl in zip(shards, md, keys, flts):
        ops.kcenter_init(feats, lo, hi, cen, m, k, fl)
    picks = []
    centre = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(K):
        allk = torch.stack(keys)
        best = allk[:, 0].max()
        row = int(torch.where(allk[:, 0] == best, allk[:, 1], torch.full_like(allk[:, 1], 2 ** 62)).min())
        picks.append(row)
        centre.fill_(row)
        for (lo, hi), m, k, fl in zip(shards, md, keys, flts):
            ops.kcenter_step(feats, lo, hi, centre, m, k, fl)
    assert picks == want.cpu().tolist()
    assert torch.equal(torch.cat(md).sqrt(), wm)


def test_kcenter_baseline_size_filtered_equals_exact():
    ops = _ops()
    N, D, L, K = 10000, 2048, 50, 500
    feats = torch.from_numpy(synth.coreset_features(11, N, D)).cuda()
    p0, m0 = ops.kcenter_greedy(feats, list(range(L)), K)
    flt = ops.KCenterFilter(feats)
    p1, m1 = ops.kcenter_greedy(feats, list(range(L)), K, flt)
    assert p0.cpu().tolist() == p1.cpu().tolist() and torch.equal(m0, m1)
    exact, screened = flt.stats()
    assert exact < 0.25 * screened


@pytest.mark.parametrize("descending", [True, False])
def test_topk_records_and_merge_equal_the_unsharded_stable_ranking(descending):
    """das_topk_records + das_topk_merge (the multi-GPU candidate exchange, here with the 'ranks' simulated by slicing):
    every shard writes k record slots (padding behind a short shard, an EMPTY shard writes padding only), the gathered
    table is merged on the device; the result must be the stable ranking of the whole pool - ties by global index."""
    from deep_active_semantic_segmentation_b200 import ops
    rng = np.random.default_rng(12)
    n, k, W = 103, 17, 5
    scores = np.round(rng.standard_normal(n), 1).astype(np.float32)     # many exact ties
    scores[7] = scores[50] = scores[90] = scores.max() if descending else scores.min()   # a tie across shards at the top
    bounds = [0, 40, 40, 41, 80, n]                                      # shard 1 is empty, shard 2 has one row (< k)
    blocks = []
    for r in range(W):
        lo, hi = bounds[r], bounds[r + 1]
        rec = ops.topk_records(torch.from_numpy(scores[lo:hi]).cuda(), k, descending, id_offset=lo)
        assert rec.shape == (k, 2) and rec.dtype == torch.int64
        s_loc, i_loc = ops.records_to_host(rec)
        assert len(i_loc) == min(k, hi - lo)
        assert i_loc.tolist() == [lo + j for j in R.rank_topk(scores[lo:hi].tolist(), k, descending)]
        np.testing.assert_array_equal(s_loc, scores[i_loc])
        blocks.append(rec)
    merged = ops.topk_merge(torch.cat(blocks), k, descending)
    ms, mi = ops.records_to_host(merged)
    assert mi.tolist() == R.rank_topk(scores.tolist(), k, descending)
    np.testing.assert_array_equal(ms, scores[mi])
    # payload ids instead of positions (region candidates: flat pool indices, -1 = unused slot with a -inf score)
    ids = torch.arange(1000, 1000 + n, dtype=torch.int64).cuda()
    sc = torch.from_numpy(scores).cuda().clone()
    sc[::9] = float("-inf")
    ids[::9] = -1
    rec = ops.topk_records(sc, 30, True, ids=ids)
    s_p, i_p = ops.records_to_host(rec)
    keep = [j for j in R.rank_topk(sc.cpu().tolist(), 30, True) if j % 9 != 0]
    assert i_p.tolist() == [1000 + j for j in keep]
    # fewer candidates than slots everywhere: the merge returns what exists, padding is dropped on the host
    small = ops.topk_merge(torch.cat([ops.topk_records(torch.from_numpy(scores[:3]).cuda(), 8, descending),
                                      ops.topk_records(torch.from_numpy(scores[3:5]).cuda(), 8, descending, id_offset=3)]), 8, descending)
    assert ops.records_to_host(small)[1].tolist() == R.rank_topk(scores[:5].tolist(), 8, descending)


def test_suppress_rects_many_host_records_and_device_records():
    """das_suppress_rects_host passes the records in the kernel parameters, 128 per launch: 300 records (three launches),
    clipped and degenerate rectangles included, must equal the Python slices of mc_dropout.py:110-121; device records
    (das_suppress_rects) give the same."""
    from deep_active_semantic_segmentation_b200 import ops
    rng = np.random.default_rng(21)
    B, H, W = 5, 37, 53
    m = rng.random((B, H, W)).astype(np.float32) + 1
    rects = [(int(rng.integers(0, B)), int(rng.integers(0, H + 5)), int(rng.integers(0, W + 5)), int(rng.integers(0, 9)),
              int(rng.integers(0, 9))) for _ in range(300)]
    want = m.copy()
    for (i, r, c, h, w) in rects:
        want[i, r:r + h, c:c + w] = 0
    t = torch.from_numpy(m.copy()).cuda()
    ops.suppress_rects(t, rects)
    np.testing.assert_array_equal(t.cpu().numpy(), want)
    t2 = torch.from_numpy(m.copy()).cuda()
    ops.suppress_rects(t2, torch.tensor(rects, dtype=torch.int32).cuda())
    np.testing.assert_array_equal(t2.cpu().numpy(), want)
    t3 = torch.from_numpy(m.copy()).cuda()
    ops.suppress_rects(t3, [])
    np.testing.assert_array_equal(t3.cpu().numpy(), m)
