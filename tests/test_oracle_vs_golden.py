"""CPU: the restatement (oracle/restate.py) must reproduce what the reference's own code
produced (tests/golden/*.npz, made by oracle/gen_golden.py) - this is what pins the oracle."""
import numpy as np
import pytest

from oracle import restate as R
from tests import golden_util as G

RTOL = 1e-5   # north_star tolerance (relative, float32)
ATOL = 1e-6   # absolute floor for per-pixel maps / means that can be ~0


def _mc_case(name, n=None):
    g = G.load(name)
    seed, N, T, C, H, W, block, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, H, W, block, g["logits_sha"], n)
    return g, (seed, N, T, C, H, W, block, k, bs), logits, labels


@pytest.mark.parametrize("name,n", [("mc_small", None), ("mc_aligned", None), ("mc_config1", 4)])
def test_vote_entropy_and_ceal_match_reference(name, n):
    g, (seed, N, T, C, H, W, block, k, bs), logits, labels = _mc_case(name, n)
    n = N if n is None else n
    sc = {key: [] for key in R.SCORE_NAMES}
    ceal = {"pred_entropy": [], "confidence": [], "margin": []}
    for i in range(n):
        maps = R.mc_maps(logits[i], labels[i], C)
        if i < g["ve_maps"].shape[0]:
            np.testing.assert_allclose(maps["vote_entropy"], g["ve_maps"][i], rtol=RTOL, atol=ATOL)
        s = R.image_scores(maps)
        for key in sc:
            sc[key].append(s[key])
        single = R.image_scores(R.mc_maps(logits[i, :1], labels[i], C))   # CEAL = the T=1 case
        for key in ceal:
            ceal[key].append(single[key])
    np.testing.assert_allclose(sc["vote_entropy"], g["ve_scores"][:n], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ceal["pred_entropy"], g["ceal_entropy"][:n], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ceal["confidence"], g["ceal_conf"][:n], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ceal["margin"], g["ceal_margin"][:n], rtol=RTOL, atol=1e-7)
    if n == N:
        assert R.rank_topk(sc["vote_entropy"], k, True) == g["ve_selected"].tolist()
        assert R.rank_topk(ceal["pred_entropy"], k, True) == g["ceal_entropy_selected"].tolist()
        assert R.rank_topk(ceal["confidence"], k, False) == g["ceal_conf_selected"].tolist()
        assert R.rank_topk(ceal["margin"], k, False) == g["ceal_margin_selected"].tolist()
        # weak labels = argmax of the deterministic pass, 255 where invalid (ceal.py:157-163)
        for j, i in enumerate(g["weak_idx"].tolist()):
            wl = R.votes_from_logits(logits[i, 0]).copy()
            wl[~R.valid_mask(labels[i], C)] = 255
            np.testing.assert_array_equal(wl, g["weak_labels"][j])


def test_golden_rankings_are_consistent_with_golden_scores():
    # the stored selections are the stable sort of the stored scores (pins rank_topk's tie rule)
    for name in ("mc_small", "mc_aligned", "mc_config1"):
        g = G.load(name)
        k = int(g["meta"][7])
        assert R.rank_topk(g["ve_scores"].tolist(), k, True) == g["ve_selected"].tolist()
        assert R.rank_topk(g["ceal_conf"].tolist(), k, False) == g["ceal_conf_selected"].tolist()


def test_rank_topk_is_stable_on_ties():
    s = [0.5, 0.25, 0.5, 0.75, 0.25, 0.5]
    assert R.rank_topk(s, 4, True) == [3, 0, 2, 5]
    assert R.rank_topk(s, 3, False) == [1, 4, 0]
    assert R.rank_topk(s, 10, True) == [3, 0, 2, 5, 1, 4]
    assert R.rank_topk([], 3, True) == []


@pytest.mark.parametrize("name", ["region_small", "region_mid"])
def test_region_selection_matches_reference(name):
    g = G.load(name)
    seed, N, T, C, S, block, Rg, sel_size, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, S, S, block, g["logits_sha"])
    existing = G.regions_from_rows(g["existing"], N)
    ve = [R.mc_maps(logits[i], labels[i], C)["vote_entropy"] for i in range(N)]
    (regions, count), norm = R.region_selection(ve, existing, Rg, sel_size, S)
    np.testing.assert_allclose(norm, g["norm_maps"], rtol=RTOL, atol=ATOL)
    assert count == int(g["count"])
    assert regions == G.regions_from_rows(g["regions"], N)
    # sharded formulation (per-image sequences + merge) == the sequential global loop
    K = float(g["K"])
    seqs = [R.nms_sequence_single(norm[i].copy(), Rg, int(np.ceil(K))) for i in range(N)]
    merged, mcount = R.merge_nms_sequences(seqs, Rg, K, norm.shape[1], norm.shape[2])
    assert (merged, mcount) == (regions, count)


def test_nms_png_fixture():
    """reference active_selection/tests.py:213-231 (the reference's own data-free NMS case)."""
    g = G.load("nms_png")
    Rg, K = int(g["R"]), int(g["K"])
    raw = np.stack([R.box_sum(im.astype(np.float32) / 256, Rg) for im in g["images"]])
    np.testing.assert_array_equal(raw, g["raw_maps"])       # k/256 sums are exact in float32
    assert raw.max() == np.float32(8860.890625)
    norm = R.minmax_normalise(raw)
    np.testing.assert_array_equal(norm, g["norm_maps"])
    regions, count = R.square_nms(norm.copy(), Rg, K)
    assert count == 10 == int(g["count"])
    assert regions == G.regions_from_rows(g["regions"], 2)
    assert regions[0] == [(18, 72, 127, 127), (275, 294, 127, 127), (114, 199, 127, 127), (362, 0, 127, 127), (241, 166, 127, 127)]
    seqs = [R.nms_sequence_single(norm[i].copy(), Rg, K) for i in range(2)]
    assert R.merge_nms_sequences(seqs, Rg, K, 386, 386) == (regions, count)


def test_nms_degenerate_all_below_threshold():
    # the first pick is unconditional, then the loop stops (mc_dropout.py:91-106)
    m = np.full((3, 5, 5), 0.001, dtype=np.float32)
    ref_regions, ref_count = R.square_nms(m.copy(), 2, 4.0)
    assert ref_count == 1 and ref_regions[0] == [(0, 0, 2, 2)]
    seqs = [R.nms_sequence_single(m[i].copy(), 2, 4) for i in range(3)]
    assert R.merge_nms_sequences(seqs, 2, 4.0, 5, 5) == (ref_regions, ref_count)


def test_nms_merge_equals_sequential_on_random_pools():
    rng = np.random.default_rng(5)
    for trial in range(12):
        N, H2, W2, Rg = int(rng.integers(1, 6)), int(rng.integers(6, 30)), int(rng.integers(6, 30)), int(rng.integers(2, 7))
        m = rng.random((N, H2, W2)).astype(np.float32)
        m[rng.random(m.shape) < 0.3] = 0
        if trial % 3 == 0:
            m = np.round(m, 1)       # many exact ties
        K = float(rng.integers(1, 40)) + 0.5
        ref = R.square_nms(m.copy(), Rg, K)
        seqs = [R.nms_sequence_single(m[i].copy(), Rg, int(np.ceil(K))) for i in range(N)]
        assert R.merge_nms_sequences(seqs, Rg, K, H2, W2) == ref


def test_noise_fixture():
    g = G.load("noise_small")
    seed, N, T, C, S, block, Rg, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, 2 * T, C, S, S, block, g["logits_sha"])
    first = [R.mc_maps(logits[i, :T], labels[i], C)["vote_entropy"] for i in range(N)]
    second = [R.mc_maps(logits[i, T:], labels[i], C)["vote_entropy"] for i in range(N)]
    s_first = [float(np.float32(m.sum(dtype=np.float64)) / (S * S)) for m in first]
    s_comb = [float(np.float32((a + b).sum(dtype=np.float64)) / (S * S)) for a, b in zip(first, second)]
    # every mc_noise scorer starts a fresh replay model -> passes 0..T-1 (mc_noise.py:46-60,116-129)
    np.testing.assert_allclose(s_first, g["input_noise_scores"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(s_first, g["feature_noise_scores"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(s_comb, g["combined_scores"], rtol=RTOL, atol=1e-7)   # mc_noise.py:141-143
    assert R.rank_topk(s_comb, k, True) == g["combined_selected"].tolist()
    existing = G.regions_from_rows(g["existing"], N)
    (regions, count), _ = R.region_selection([a + b for a, b in zip(first, second)], existing, Rg, 1, S)
    assert count == int(g["count"]) and regions == G.regions_from_rows(g["regions"], N)


def test_kcenter_toy_fixture():
    g = G.load("kcenter_toy")
    picks, m = R.kcenter_greedy(g["features"], g["selected"].tolist(), 5)
    assert picks == [0, 2, 8, 4, 7] == g["picks"].tolist()
    assert abs(m.max() - 1.41421) < 1e-5


@pytest.mark.parametrize("name", ["coreset_small", "coreset_mid"])
def test_kcenter_matches_reference(name):
    from deep_active_semantic_segmentation_b200 import synth

    g = G.load(name)
    seed, N, D, L, K = (int(v) for v in g["meta"])
    feats = synth.coreset_features(seed, N, D)
    assert G.sha(feats) == str(g["features_sha"])
    picks, m = R.kcenter_greedy(feats, list(range(L)), K)
    assert picks == g["picks"].tolist()
    # d = sqrt(|x|^2+|y|^2-2x.y) cancels to ~sqrt(eps64*|x|^2) ~ 1e-5 of noise where the true distance is 0
    np.testing.assert_allclose(m, g["min_dist"], rtol=1e-9, atol=1e-4)


def test_kcenter_restatement_matches_reference_at_baseline_size():
    """BASELINE config 5 (N = 10 000, D = 2048, L = 50, K = 500): 500 picks of the reference's sklearn loop."""
    from deep_active_semantic_segmentation_b200 import synth

    g = G.load("coreset_baseline")
    seed, N, D, L, K = (int(v) for v in g["meta"])
    feats = synth.coreset_features(seed, N, D)
    assert G.sha(feats) == str(g["features_sha"])
    picks, m = R.kcenter_greedy(feats, list(range(L)), K)
    assert picks == g["picks"].tolist()
    np.testing.assert_allclose(m.max(), float(g["min_dist_max"]), rtol=1e-9)
    np.testing.assert_allclose(m[:256], g["min_dist_head"], rtol=1e-9, atol=1e-4)


def test_restatement_matches_reference_at_config2_size():
    """BASELINE config 2 plane (512 x 1024, C = 19, T = 20), image 0 of the mc_baseline golden: reference vote entropy
    (score + sampled pixels) and the reference's single-pass CEAL scores."""
    from deep_active_semantic_segmentation_b200 import synth

    g = G.load("mc_baseline")
    seed, N, T, C, H, W, block, bs = (int(v) for v in g["meta"])
    logits = synth.pool_logits(seed, [0], T, C, H, W, block)[0]
    labels = synth.pool_labels(seed, [0], H, W, C, block)[0]
    valid = R.valid_mask(labels, C)
    ve = R.vote_entropy_map(R.votes_from_logits(logits), C, valid)
    rows, cols = g["px_rows"].astype(np.int64), g["px_cols"].astype(np.int64)
    np.testing.assert_allclose(ve[rows, cols], g["ve_px"][0], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.float32(ve.mean(dtype=np.float64)), g["ve_scores"][0], rtol=RTOL, atol=1e-7)
    single = R.image_scores(R.mc_maps(logits[:1], labels, C))
    np.testing.assert_allclose(single["pred_entropy"], g["ceal_entropy"][0], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(single["confidence"], g["ceal_conf"][0], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(single["margin"], g["ceal_margin"][0], rtol=RTOL, atol=1e-7)


def test_kcenter_asserts_on_reselection():
    feats = np.zeros((4, 3), dtype=np.float32)     # all distances 0 -> argmax = 0, already selected
    with pytest.raises(AssertionError):
        R.kcenter_greedy(feats, [0], 1)


def test_coreset_e2e_fixture():
    g = G.load("coreset_e2e")
    seed, N, L, K, bs = (int(v) for v in g["meta"])
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    feats = rng.standard_normal(size=(N, 128, 64, 64), dtype=np.float32)
    feats += (rng.integers(0, 4, size=(N, 1, 1, 1)) * 0.5).astype(np.float32)
    assert G.sha(feats) == str(g["features_sha"])
    rows = np.stack([R.avg_pool_features(feats[i], 32) for i in range(N)])
    assert rows.shape == (N, 1152)
    picks, _ = R.kcenter_greedy(rows, list(range(L)), K)
    assert picks == g["chosen"].tolist()      # combined = already + candidates -> index == path


def test_cpu_port_matches_restatement():
    """oracle/cpu_port.py (the timed CPU baseline of bench.py) computes the same scores."""
    import torch
    from deep_active_semantic_segmentation_b200 import synth
    from oracle import cpu_port

    B, T, C, H, W = 2, 6, 19, 24, 40
    logits = synth.pool_logits(5, list(range(B)), T, C, H, W, 8)
    labels = synth.pool_labels(5, list(range(B)), H, W, C, 8)
    got = cpu_port.score_batch([torch.from_numpy(np.ascontiguousarray(logits[:, t])) for t in range(T)],
                               torch.from_numpy(labels), C)
    for b in range(B):
        want = R.image_scores(R.mc_maps(logits[b], labels[b], C))
        for k in got:
            np.testing.assert_allclose(got[k][b].item(), want[k], rtol=RTOL, atol=1e-6, err_msg=k)


def _unet_head(seed, N, S):
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 911]))
    coarse = rng.standard_normal(size=(N, 2, -(-S // 8), -(-S // 8)), dtype=np.float32) * np.float32(2.0)
    up = np.repeat(np.repeat(coarse, 8, axis=2), 8, axis=3)[:, :, :S, :S]
    return (up + rng.standard_normal(size=(N, 2, S, S), dtype=np.float32)).astype(np.float32)


def test_accuracy_selectors_match_reference():
    """oracle accuracy_scores / accuracy_error_map vs what the reference's ActiveSelectionAccuracy ranked."""
    g = G.load("accuracy_small")
    seed, N, C, S, block, Rg, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, 1, C, S, S, block, g["logits_sha"])
    unet = _unet_head(seed, N, S)
    assert G.sha(unet) == str(g["unet_sha"])
    sc = [R.accuracy_scores(logits[i, 0], unet[i], labels[i], C) for i in range(N)]
    for key, col in (("labels", "wrong_count"), ("softmax", "p0_sum"), ("argmax", "not_argmax_sum"), ("unsure", "unsure_mean")):
        mine = np.array([s[col] for s in sc], dtype=np.float32)
        np.testing.assert_allclose(mine, g[key + "_scores"], rtol=1e-5, atol=1e-6, err_msg=key)
        assert R.rank_topk(g[key + "_scores"].tolist(), k, True) == g[key + "_selected"].tolist()
    maps = np.stack([R.accuracy_error_map(unet[i], labels[i], C) for i in range(N)])
    existing = G.regions_from_rows(g["existing"], N)
    (regions, count), _ = R.region_selection(maps, existing, Rg, 1, S)
    assert count == int(g["count"])
    assert regions == G.regions_from_rows(g["regions"], N)


def test_maxsubset_matches_reference():
    """oracle max_representative_samples vs the reference class on its own seeded fixture (tests.py:616-645) and on
    two float32 pools with duplicated candidates."""
    g = G.load("maxsubset")
    np.random.seed(seed=int(g["ref_seed"]))
    images = np.concatenate((np.random.normal(loc=2.0, scale=1.0, size=(400, 1024)),
                             np.random.normal(loc=4.0, scale=1.0, size=(400, 1024)),
                             np.random.normal(loc=6.0, scale=1.0, size=(150, 1024)),
                             np.random.normal(loc=4.0, scale=3.0, size=(50, 1024))), axis=0)
    assert G.sha(images) == str(g["ref_images_sha"])
    cand = g["ref_candidates"].tolist()
    assert R.max_representative_samples(images, images[cand, :], 4) == g["ref_picks"].tolist()
    for tag in ("a", "b"):
        X, Y, k = _maxsubset_inputs(g, tag)
        assert R.max_representative_samples(X, Y, k) == g[f"{tag}_picks"].tolist()


def _maxsubset_inputs(g, tag):
    from deep_active_semantic_segmentation_b200 import synth
    seed, N, M, D, k = (int(v) for v in g[f"{tag}_meta"])
    X = synth.coreset_features(seed, N, D)
    rng = np.random.default_rng(seed)
    ci = rng.integers(0, N, size=M)
    Y = (X[ci] + np.float32(0.05) * rng.standard_normal((M, D)).astype(np.float32)).astype(np.float32)
    Y[M // 3] = Y[M // 5]
    assert G.sha(X) == str(g[f"{tag}_x_sha"]) and G.sha(Y) == str(g[f"{tag}_y_sha"])
    return X, Y, k


# ---- fused final upsample: models/deeplab.py:59 restated ------------------------

UPSAMPLE_CASES = ("upsample_odd", "upsample_rect", "upsample_mid")


@pytest.mark.parametrize("name", UPSAMPLE_CASES)
def test_bilinear_restatement_matches_aten(name):
    g, m, low, _ = G.upsample_case(name)
    C = m["C"]
    got = R.bilinear_upsample_align_corners(low[0, 0][[0, C // 2, C - 1]], m["H"], m["W"])
    want = g["upsampled_image0_pass0_classes"]
    if name == "upsample_rect":
        # 12x16 -> 48x64: ATen rounds this small even-sized case differently (same weights, other association)
        np.testing.assert_allclose(got, want, rtol=0, atol=2 * np.spacing(np.float32(np.abs(want).max())))
    else:
        np.testing.assert_array_equal(got, want)   # bit-exact at the DeepLab-style shapes


@pytest.mark.parametrize("name", UPSAMPLE_CASES)
def test_upsampled_scoring_matches_reference(name):
    """reference selectors on F.interpolate(low_res_x) == restatement on bilinear_upsample_align_corners(low_res_x)"""
    g, m, low, labels = G.upsample_case(name)
    N, C, k = m["N"], m["C"], m["k"]
    ve, ent, conf, marg = [], [], [], []
    for i in range(N):
        maps = R.mc_maps_upsampled(low[i], labels[i], C, m["H"], m["W"])
        if i < g["ve_maps"].shape[0]:
            # a vote can flip where two upsampled logits tie within an ulp: tolerate a handful of pixels
            bad = ~np.isclose(maps["vote_entropy"], g["ve_maps"][i], rtol=RTOL, atol=ATOL)
            assert bad.sum() <= (2 if name == "upsample_rect" else 0), bad.sum()
        ve.append(R.image_scores(maps)["vote_entropy"])
        single = R.image_scores(R.mc_maps_upsampled(low[i, :1], labels[i], C, m["H"], m["W"]))
        ent.append(single["pred_entropy"]), conf.append(single["confidence"]), marg.append(single["margin"])
    vtol = 1e-5 if name == "upsample_rect" else 1e-7
    np.testing.assert_allclose(ve, g["ve_scores"], rtol=RTOL, atol=vtol)
    np.testing.assert_allclose(ent, g["ceal_entropy"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(conf, g["ceal_conf"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(marg, g["ceal_margin"], rtol=RTOL, atol=1e-7)
    assert R.rank_topk(ve, k, True) == g["ve_selected"].tolist()
    assert R.rank_topk(ent, k, True) == g["ceal_entropy_selected"].tolist()
    assert R.rank_topk(conf, k, False) == g["ceal_conf_selected"].tolist()
    assert R.rank_topk(marg, k, False) == g["ceal_margin_selected"].tolist()


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/models"), reason="needs the reference tree (build container only)")
def test_restatement_on_the_reference_deeplab_mobilenet_with_live_dropout():
    """BASELINE config 1 as written: the reference's own DataParallel-style DeepLab-MobileNet (pretrained=False, seed 0),
    Dropout2d live (mc_dropout.py:175-178), T = 5, 513 x 513, 21 classes, through the reference's selector on the CPU.
    The logits of every pass are captured with a forward hook; the restatement on exactly those logits must give the
    reference's scores and ranking (2 images keep it to ~10 s of CPU)."""
    import torch
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    from models.deeplab import DeepLab          # reference model, unmodified

    N, T, C, S, bs = 2, 5, 21, 513, 2
    torch.manual_seed(0)
    net = DeepLab(num_classes=C, backbone="mobilenet", output_stride=16, sync_bn=False, freeze_bn=False, pretrained=False)
    net.eval()
    seen = []
    net.register_forward_hook(lambda m, i, o: seen.append(o.detach().clone()))
    g = torch.Generator().manual_seed(1)
    images = torch.randn((N, 3, S, S), generator=g)
    from deep_active_semantic_segmentation_b200 import synth
    labels = synth.pool_labels(9, list(range(N)), S, S, C, 32)

    class DS:
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            j = int(self.paths[i])
            return {"image": images[j], "label": torch.from_numpy(labels[j])}

    class Wrapper(torch.nn.Module):             # what nn.DataParallel looks like to the selector: .module + call
        def __init__(self, m):
            super().__init__()
            self.module = m

        def forward(self, x):
            return self.module(x)

    old_ds, old_T = ref.paths_dataset.PathsDataset, ref.constants.MC_STEPS
    ref.paths_dataset.PathsDataset, ref.constants.MC_STEPS = DS, T
    captured = {}
    import builtins
    real_sorted = builtins.sorted

    def spy(iterable, key=None, reverse=False):
        items = list(iterable)
        captured["scores"] = [float(x[0]) for x in items]
        return real_sorted(items, key=key, reverse=reverse)

    ref.mc_dropout.sorted = spy
    try:
        sel = ref.active_selection.get_active_selection_class("variance", C, None, S, bs)
        chosen = sel.get_vote_entropy_for_images(Wrapper(net), [str(i) for i in range(N)], N)
    finally:
        del ref.mc_dropout.sorted
        ref.paths_dataset.PathsDataset, ref.constants.MC_STEPS = old_ds, old_T
    assert len(seen) == T and seen[0].shape == (N, C, S, S)
    assert not torch.equal(seen[0], seen[1])                                  # the decoder / ASPP Dropout2d were live
    stack = torch.stack(seen, dim=1).numpy()                                  # [N,T,C,S,S]
    mine = [R.image_scores(R.mc_maps(stack[i], labels[i], C))["vote_entropy"] for i in range(N)]
    np.testing.assert_allclose(mine, captured["scores"], rtol=RTOL, atol=1e-7)
    assert [int(p) for p in chosen] == R.rank_topk(mine, N, True)
    assert not any(m.training for m in net.modules() if isinstance(m, torch.nn.Dropout2d))   # model.eval() restored
