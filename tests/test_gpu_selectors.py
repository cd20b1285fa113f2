"""GPU: the selector mirror (same method signatures as the reference's active_selection classes) must
reproduce what the reference's own classes returned on the same synthetic pools (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from deep_active_semantic_segmentation_b200 import synth
from tests import fakes
from tests import golden_util as G

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-7


@pytest.fixture(autouse=True)
def _synthetic_data_layer():
    from deep_active_semantic_segmentation_b200.active_selection import base
    from deep_active_semantic_segmentation_b200 import constants
    old, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
    base.paths_dataset.PathsDataset = fakes.SyntheticPathsDataset
    yield
    base.paths_dataset.PathsDataset, constants.MC_STEPS = old, old_T


def _set_T(T):
    import sys
    from deep_active_semantic_segmentation_b200 import constants
    constants.MC_STEPS = T
    if "constants" in sys.modules and hasattr(sys.modules["constants"], "MC_STEPS"):
        sys.modules["constants"].MC_STEPS = T      # honoured like the reference's global (constants.py:6)


def _factory(method, C, pool, crop, bs):
    from deep_active_semantic_segmentation_b200.active_selection import get_active_selection_class
    return get_active_selection_class(method, C, pool, crop, bs)


def _paths(N):
    return [str(i) for i in range(N)]


def _idx(paths):
    return [int(p) for p in paths]


@pytest.mark.parametrize("name", ["mc_small", "mc_aligned", "mc_config1"])
@pytest.mark.parametrize("pass_group", [1, 4])
def test_mc_dropout_and_ceal_selectors_match_reference(name, pass_group):
    g = G.load(name)
    seed, N, T, C, H, W, block, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, H, W, block, g["logits_sha"])
    pool = fakes.Pool(logits, labels)
    crop = H if H == W else -1
    _set_T(T)

    sel = _factory("variance", C, pool, crop, bs)
    sel.pass_group = pass_group
    model = fakes.ReplayModel(pool)
    chosen = sel.get_vote_entropy_for_images(model, _paths(N), k)
    assert isinstance(chosen, tuple) and _idx(chosen) == g["ve_selected"].tolist()
    np.testing.assert_allclose(sel.last_scores, g["ve_scores"], rtol=RTOL, atol=ATOL)
    assert not model.drop.training and model.dropout_train_calls == sum(model.calls.values()) == T * -(-N // bs)

    # per-pixel maps of the first batch through the reference's private method signature
    nb = min(bs, N)
    ds = fakes.SyntheticPathsDataset(pool, _paths(nb), crop, include_labels=True)
    ib = torch.stack([ds[i]["image"] for i in range(nb)]).cuda()
    lb = torch.stack([ds[i]["label"] for i in range(nb)]).cuda()
    maps = sel._get_vote_entropy_for_batch(fakes.ReplayModel(pool), ib, lb)
    assert len(maps) == nb and maps[0].shape == (H, W)
    np.testing.assert_allclose(torch.stack(maps).cpu().numpy(), g["ve_maps"], rtol=RTOL, atol=2e-6)

    ceal = _factory("ceal_entropy", C, pool, crop, bs)
    ent_sel, ent = ceal.get_maximum_entropy_samples(fakes.ReplayModel(pool), _paths(N), k)
    assert _idx(ent_sel) == g["ceal_entropy_selected"].tolist()
    np.testing.assert_allclose(ent, g["ceal_entropy"], rtol=RTOL, atol=ATOL)
    conf_sel = ceal.get_least_confident_samples(fakes.ReplayModel(pool), _paths(N), k)
    assert _idx(conf_sel) == g["ceal_conf_selected"].tolist()
    np.testing.assert_allclose(ceal.last_scores, g["ceal_conf"], rtol=RTOL, atol=ATOL)
    marg_sel = ceal.get_least_margin_samples(fakes.ReplayModel(pool), _paths(N), k)
    assert _idx(marg_sel) == g["ceal_margin_selected"].tolist()
    np.testing.assert_allclose(ceal.last_scores, g["ceal_margin"], rtol=RTOL, atol=ATOL)
    fused = ceal.get_fusion_of_confidence_margin_entropy_samples(fakes.ReplayModel(pool), _paths(N), k)
    assert len(fused) == min(k, N) and set(fused) <= set(ent_sel) | set(conf_sel) | set(marg_sel)
    weak = ceal.get_weakly_labeled_data(fakes.ReplayModel(pool), _paths(N), float(g["weak_threshold"]), entropies=list(g["ceal_entropy"]))
    assert _idx(weak.keys()) == g["weak_idx"].tolist()
    for j, p in enumerate(weak):
        assert weak[p].dtype == np.uint8
        np.testing.assert_array_equal(weak[p], g["weak_labels"][j])


@pytest.mark.parametrize("name", ["upsample_odd", "upsample_rect", "upsample_mid"])
def test_selectors_on_low_resolution_logits_match_reference(name):
    """the model hands over `low_res_x` (models/deeplab.py:58); the goldens are the reference
    selectors on F.interpolate(low_res_x) (models/deeplab.py:59).  Same selector calls, same results."""
    g, m, low, labels = G.upsample_case(name)
    N, T, C, k, bs = m["N"], m["T"], m["C"], m["k"], m["batch_size"]
    pool = fakes.LowResPool(low, labels, m["H"], m["W"])
    crop = m["H"] if m["H"] == m["W"] else -1
    _set_T(T)
    from deep_active_semantic_segmentation_b200 import _lib
    n0 = _lib.launch_count()
    sel = _factory("variance", C, pool, crop, bs)
    model = fakes.ReplayModel(pool)
    chosen = sel.get_vote_entropy_for_images(model, _paths(N), k)
    assert _idx(chosen) == g["ve_selected"].tolist() and not model.drop.training
    vtol = 1e-5 if name == "upsample_rect" else ATOL      # ATen's small even-sized path rounds differently: a vote may flip
    np.testing.assert_allclose(sel.last_scores, g["ve_scores"], rtol=RTOL, atol=vtol)
    nb = -(-N // bs)
    assert 2 * nb < _lib.launch_count() - n0 <= 2 * nb + 3   # ONE fused-upsample launch + one reduce per batch, then the top-k
    ceal = _factory("ceal_entropy", C, pool, crop, bs)
    ent_sel, ent = ceal.get_maximum_entropy_samples(fakes.ReplayModel(pool), _paths(N), k)
    assert _idx(ent_sel) == g["ceal_entropy_selected"].tolist()
    np.testing.assert_allclose(ent, g["ceal_entropy"], rtol=RTOL, atol=ATOL)
    marg_sel = ceal.get_least_margin_samples(fakes.ReplayModel(pool), _paths(N), k)
    assert _idx(marg_sel) == g["ceal_margin_selected"].tolist()
    np.testing.assert_allclose(ceal.last_scores, g["ceal_margin"], rtol=RTOL, atol=ATOL)
    # a factor the fused kernel does not take is interpolated by torch on the device and scored as usual
    big = fakes.LowResPool(low, labels[:, :2 * m["h"] - 1, :2 * m["w"] - 1], 2 * m["h"] - 1, 2 * m["w"] - 1)
    sel2 = _factory("variance", C, big, -1, bs)
    out = sel2.get_vote_entropy_for_images(fakes.ReplayModel(big), _paths(N), k)
    assert len(out) == min(k, N) and all(np.isfinite(sel2.last_scores))


@pytest.mark.parametrize("name", ["region_small", "region_mid"])
def test_region_selector_matches_reference(name):
    g = G.load(name)
    seed, N, T, C, S, block, Rg, sel_size, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, S, S, block, g["logits_sha"])
    pool = fakes.Pool(logits, labels)
    _set_T(T)
    sel = _factory("variance", C, pool, S, bs)
    model = fakes.ReplayModel(pool)
    regions, count = sel.create_region_maps(model, _paths(N), G.regions_from_rows(g["existing"], N), Rg, sel_size)
    assert count == int(g["count"])
    want = {str(i): lst for i, lst in enumerate(G.regions_from_rows(g["regions"], N)) if lst}
    assert regions == want                      # dict path -> [(r,c,R,R)], pick order preserved
    assert not model.drop.training


def test_mc_noise_selectors_match_reference():
    g = G.load("noise_small")
    seed, N, T, C, S, block, Rg, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, 2 * T, C, S, S, block, g["logits_sha"])
    pool = fakes.Pool(logits, labels)
    _set_T(T)
    sel = _factory("noise_variance", C, pool, S, bs)
    m = fakes.ReplayModel(pool)
    assert _idx(sel.get_vote_entropy_for_images_with_input_noise(m, _paths(N), k)) == g["input_noise_selected"].tolist()
    np.testing.assert_allclose(sel.last_scores, g["input_noise_scores"], rtol=RTOL, atol=ATOL)
    m = fakes.ReplayModel(pool)
    assert _idx(sel.get_vote_entropy_for_images_with_feature_noise(m, _paths(N), k)) == g["feature_noise_selected"].tolist()
    np.testing.assert_allclose(sel.last_scores, g["feature_noise_scores"], rtol=RTOL, atol=ATOL)
    assert m.noisy_calls == sum(m.calls.values()) and not m.noisy_features      # flag set for every pass, then cleared
    m = fakes.ReplayModel(pool)
    assert _idx(sel.get_vote_entropy_for_batch_with_noise_and_vote_entropy(m, _paths(N), k)) == g["combined_selected"].tolist()
    np.testing.assert_allclose(sel.last_scores, g["combined_scores"], rtol=RTOL, atol=ATOL)
    assert m.noisy_calls == m.dropout_train_calls == sum(m.calls.values()) // 2
    regions, count = sel.create_region_maps(fakes.ReplayModel(pool), _paths(N), G.regions_from_rows(g["existing"], N), Rg, 1)
    assert count == int(g["count"])
    assert regions == {str(i): lst for i, lst in enumerate(G.regions_from_rows(g["regions"], N)) if lst}


def test_coreset_selector_matches_reference():
    g = G.load("coreset_e2e")
    seed, N, L, K, bs = (int(v) for v in g["meta"])
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    feats = rng.standard_normal(size=(N, 128, 64, 64), dtype=np.float32)
    feats += (rng.integers(0, 4, size=(N, 1, 1, 1)) * 0.5).astype(np.float32)
    assert G.sha(feats) == str(g["features_sha"])
    pool = fakes.Pool(np.zeros((N, 1, 2, 64, 64), np.float32), None, feats)
    sel = _factory("coreset", 2, pool, 64, bs)
    model = fakes.ReplayModel(pool, "enet")
    chosen = sel.get_k_center_greedy_selections(K, model, _paths(N)[L:], _paths(N)[:L])
    assert _idx(chosen) == g["chosen"].tolist()
    assert model.return_features is False


def test_composed_mc_scores_and_factory_errors():
    from deep_active_semantic_segmentation_b200 import synth
    from oracle import restate as R
    N, T, C, H, W = 5, 6, 19, 24, 40
    logits = synth.pool_logits(3, list(range(N)), T, C, H, W, 8)
    labels = synth.pool_labels(3, list(range(N)), H, W, C, 8)
    pool = fakes.Pool(logits, labels)
    _set_T(T)
    sel = _factory("variance", C, pool, -1, 2)
    chosen, allv = sel.get_mc_scores_for_images(fakes.ReplayModel(pool), _paths(N), 3, score="bald")
    want = {k: [] for k in R.SCORE_NAMES}
    for i in range(N):
        s = R.image_scores(R.mc_maps(logits[i], labels[i], C))
        for k in want:
            want[k].append(float(s[k]))
    for k in want:
        np.testing.assert_allclose(allv[k], want[k], rtol=RTOL, atol=1e-6, err_msg=k)
    assert _idx(chosen) == R.rank_topk(want["bald"], 3, True)
    with pytest.raises(NotImplementedError):
        _factory("no_such_method", C, pool, -1, 2)
    with pytest.raises(IndexError):
        sel.get_vote_entropy_for_images(fakes.ReplayModel(pool), [], 3)       # the reference fails the same way


def test_accuracy_selectors_match_reference():
    g = G.load("accuracy_small")
    seed, N, C, S, block, Rg, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, 1, C, S, S, block, g["logits_sha"])
    rng = np.random.Generator(np.random.Philox(key=[int(seed), 911]))
    coarse = rng.standard_normal(size=(N, 2, -(-S // 8), -(-S // 8)), dtype=np.float32) * np.float32(2.0)
    unet = (np.repeat(np.repeat(coarse, 8, axis=2), 8, axis=3)[:, :, :S, :S]
            + rng.standard_normal(size=(N, 2, S, S), dtype=np.float32)).astype(np.float32)
    assert G.sha(unet) == str(g["unet_sha"])
    pool = fakes.Pool(logits, labels)

    class PairModel(torch.nn.Module):
        def __init__(self, pair):
            super().__init__()
            self.pair = pair

        @property
        def module(self):
            return self

        def forward(self, x):
            gs = [int(round(float(v) / fakes.GID_SCALE)) for v in x[:, 0, 0, 0].cpu()]
            seg = torch.from_numpy(np.stack([logits[i, 0] for i in gs])).cuda()
            return (seg, torch.from_numpy(np.stack([unet[i] for i in gs])).cuda()) if self.pair else seg

    sel = _factory("accuracy_labels", C, pool, S, bs)
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionAccuracy
    assert isinstance(sel, ActiveSelectionAccuracy) and isinstance(_factory("accuracy_eval", C, pool, S, bs), ActiveSelectionAccuracy)
    calls = (("labels", lambda: sel.get_least_accurate_sample_using_labels(PairModel(False), _paths(N), k)),
             ("softmax", lambda: sel.get_least_accurate_samples(PairModel(True), _paths(N), k, mode='softmax')),
             ("argmax", lambda: sel.get_least_accurate_samples(PairModel(True), _paths(N), k, mode='argmax')),
             ("unsure", lambda: sel.get_unsure_samples(PairModel(True), _paths(N), k)))
    for key, call in calls:
        chosen = call()
        assert isinstance(chosen, tuple) and _idx(chosen) == g[key + "_selected"].tolist(), key
        np.testing.assert_allclose(sel.last_scores, g[key + "_scores"], rtol=RTOL, atol=1e-6, err_msg=key)
    regions, count = sel.get_least_accurate_region_maps(PairModel(True), _paths(N), G.regions_from_rows(g["existing"], N), Rg, 1)
    assert count == int(g["count"])
    assert regions == {str(i): lst for i, lst in enumerate(G.regions_from_rows(g["regions"], N)) if lst}
    with pytest.raises(NotImplementedError):
        sel.get_least_accurate_samples(PairModel(True), _paths(N), k, mode='other')
    # no valid pixel at all -> the 'unsure' mean is NaN like torch's mean of an empty selection
    from deep_active_semantic_segmentation_b200 import ops
    sc = ops.accuracy_scores(torch.from_numpy(unet[:1]).cuda(), torch.full((1, S, S), 255.0).cuda(), C)
    assert np.isnan(sc[0, 3].item()) and sc[0, 4].item() == 0 and sc[0, 1].item() == 0


def test_maxsubset_selector_matches_reference():
    from deep_active_semantic_segmentation_b200.active_selection import get_max_subset_active_selector
    from tests.test_oracle_vs_golden import _maxsubset_inputs
    from oracle import restate as R
    g = G.load("maxsubset")
    sel = get_max_subset_active_selector(None, None, None)
    np.random.seed(seed=int(g["ref_seed"]))
    images = np.concatenate((np.random.normal(loc=2.0, scale=1.0, size=(400, 1024)),
                             np.random.normal(loc=4.0, scale=1.0, size=(400, 1024)),
                             np.random.normal(loc=6.0, scale=1.0, size=(150, 1024)),
                             np.random.normal(loc=4.0, scale=3.0, size=(50, 1024))), axis=0)
    cand = g["ref_candidates"].tolist()
    # the reference's own seeded fixture, float64 lists of vectors exactly as tests.py:645 passes them
    assert sel._max_representative_samples(list(images), list(images[cand, :]), 4) == g["ref_picks"].tolist()
    for tag in ("a", "b"):
        X, Y, k = _maxsubset_inputs(g, tag)
        assert sel._max_representative_samples(list(X), list(Y), k) == g[f"{tag}_picks"].tolist()
    # more picks than candidates: the reference appends None once everything is taken
    out = sel._max_representative_samples(images[:50], images[[3, 7, 7]], 4)
    assert out[:3] == R.max_representative_samples(images[:50], images[[3, 7, 7]], 3) and out[3] is None


def test_maxsubset_region_scale_properties():
    """Region-scale pool (N = 20000 cells, M = 1500 candidates, k = 750): greedy invariants that do not need the
    O(k M N) host loop - picks are distinct, and the objective after each pick is the minimum over all candidates."""
    from deep_active_semantic_segmentation_b200 import ops, synth
    X = torch.from_numpy(synth.coreset_features(3, 20000, 64)).cuda()
    rng = np.random.default_rng(0)
    Y = X[torch.from_numpy(rng.integers(0, 20000, 1500)).cuda()] + 0.01
    picks = ops.maxsubset_greedy(X, Y, 750).cpu().tolist()
    assert len(set(picks)) == 750 and min(picks) >= 0
    D = torch.cdist(X.double(), Y.double())                          # [N, M]
    md = torch.full((20000,), float("inf"), dtype=torch.float64, device="cuda")
    for step, p in enumerate(picks[:40]):
        scores = torch.minimum(md[:, None], D).sum(0)
        scores[picks[:step]] = float("inf")
        best = float(scores.min())
        assert abs(float(scores[p]) - best) <= 1e-9 * abs(best)
        md = torch.minimum(md, D[:, p])


def test_device_batch_loader_yields_what_the_dataloader_yields():
    """prefetch.DeviceBatchLoader: same batches, same order as DataLoader(shuffle=False) - on the device."""
    from torch.utils.data import DataLoader
    from deep_active_semantic_segmentation_b200.prefetch import DeviceBatchLoader

    class DictSet(torch.utils.data.Dataset):
        def __len__(self):
            return 11

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(i)
            return {"image": torch.randn(3, 9, 7, generator=g), "label": np.full((9, 7), float(i), np.float32)}

    class BareSet(DictSet):
        def __getitem__(self, i):
            return super().__getitem__(i)["image"]

    class OddSet(DictSet):        # items the fast path does not take: falls back to the DataLoader
        def __getitem__(self, i):
            return {"image": torch.zeros(2), "name": f"img{i}"}

    for ds, bs in ((DictSet(), 4), (DictSet(), 11), (DictSet(), 16), (BareSet(), 3)):
        for _ in range(2):                               # second round: cached staging buffers
            got = [({k: v.clone() for k, v in b.items()} if isinstance(b, dict) else b.clone()) for b in DeviceBatchLoader(ds, bs)]
            want = list(DataLoader(ds, batch_size=bs, shuffle=False, num_workers=0))
            assert len(got) == len(want) == len(DeviceBatchLoader(ds, bs))
            for g_, w_ in zip(got, want):
                if isinstance(w_, dict):
                    assert set(g_) == set(w_)
                    for k in w_:
                        assert g_[k].is_cuda and g_[k].dtype == w_[k].dtype
                        torch.testing.assert_close(g_[k].cpu(), w_[k], rtol=0, atol=0)
                else:
                    assert g_.is_cuda
                    torch.testing.assert_close(g_.cpu(), w_, rtol=0, atol=0)
    odd = list(DeviceBatchLoader(OddSet(), 4))
    assert len(odd) == 3 and odd[0]["name"] == ["img0", "img1", "img2", "img3"]

    # items that live in pinned memory travel without a host copy (one async copy per item and field) - same batches
    store = [{"image": torch.randn(3, 9, 7, generator=torch.Generator().manual_seed(i)).pin_memory(),
              "label": torch.full((9, 7), float(i)).pin_memory()} for i in range(11)]

    class PinnedSet(DictSet):
        def __getitem__(self, i):
            return store[i]

    class MixedSet(DictSet):      # first item pinned, later ones pageable: staged item by item, still correct
        def __getitem__(self, i):
            return store[i] if i == 0 else {k: v.clone() for k, v in store[i].items()}

    for ds, path in ((PinnedSet(), "direct"), (MixedSet(), "direct"), (DictSet(), "staged")):
        loader = DeviceBatchLoader(ds, 4)
        got = [{k: v.clone() for k, v in b.items()} for b in loader]
        assert loader.path == path and loader.batches == 3 and loader.host_seconds > 0
        want = list(DataLoader(ds, batch_size=4, shuffle=False, num_workers=0))
        for g_, w_ in zip(got, want):
            for k in w_:
                torch.testing.assert_close(g_[k].cpu(), w_[k], rtol=0, atol=0)


def test_region_selector_on_low_resolution_logits_equals_full_resolution():
    """create_region_maps with a model that returns low_res_x == create_region_maps on F.interpolate(low_res_x)
    (ATen on the CPU, which the fused kernel reproduces bit for bit at this shape: 17 -> 65)."""
    from deep_active_semantic_segmentation_b200 import synth
    N, T, C, h, H, R, bs = 5, 4, 19, 17, 65, 17, 2
    low = synth.pool_logits(31, list(range(N)), T, C, h, h, 2)
    labels = synth.pool_labels(31, list(range(N)), H, H, C, 8)
    full = torch.nn.functional.interpolate(torch.from_numpy(low).reshape(N * T, C, h, h), size=(H, H), mode="bilinear",
                                           align_corners=True).reshape(N, T, C, H, H).numpy()
    existing = [[], [(3, 5, R, R)], [], [(0, 0, H, H)], [(20, 30, R, R), (40, 2, R, R)]]
    _set_T(T)
    out = []
    for pool in (fakes.LowResPool(low, labels, H, H), fakes.Pool(full, labels)):
        sel = _factory("variance", C, pool, H, bs)
        out.append(sel.create_region_maps(fakes.ReplayModel(pool), _paths(N), existing, R, 2))
    assert out[0] == out[1] and out[0][1] > 0


@pytest.mark.parametrize("name", ["maxsubset_poolers", "maxsubset_poolers_rect"])
def test_maxsubset_feature_poolers_match_reference(name):
    """max_subset.py:49-70, 72-86, 88-111 + get_representative_regions / _images: goldens from the reference class (its
    own loops; only the avg_pool2d call that torch 2.x rejects runs with the kernel clipped to the crop - see
    oracle/gen_golden.py:gen_maxsubset_poolers)."""
    from deep_active_semantic_segmentation_b200.active_selection import get_max_subset_active_selector
    g = G.load(name)
    seed, N, F_, fh, fw, crop, region_size, bs = (int(v) for v in g["meta"])
    rng = np.random.Generator(np.random.Philox(key=[seed, 78]))
    feats = rng.standard_normal(size=(N, F_, fh, fw), dtype=np.float32)
    feats += (rng.integers(0, 3, size=(N, 1, 1, 1)) * 0.75).astype(np.float32)
    assert G.sha(feats) == str(g["features_sha"])
    pool = fakes.Pool(np.zeros((N, 1, 2, crop, crop), np.float32), None, feats)
    sel = get_max_subset_active_selector(pool, crop, bs)
    paths = _paths(N)
    model = lambda: fakes.ReplayModel(pool, "deeplab")
    cells = sel._get_features_for_image_regions(model(), paths, region_size)
    np.testing.assert_allclose(cells.cpu().numpy(), g["cell_features"], rtol=1e-5, atol=1e-6)
    cand = {}
    for row in g["cand_rows"].tolist():
        cand.setdefault(str(row[0]), []).append(tuple(row[1:]))
    li, lr = sel._convert_regions_to_list(cand)
    regs = sel._get_features_for_regions(model(), li, lr)
    np.testing.assert_allclose(regs.cpu().numpy(), g["region_features"], rtol=1e-5, atol=1e-6)
    selected, n_sel = sel.get_representative_regions(model(), paths, cand, region_size)
    assert n_sel == int(g["n_selected"])
    assert sorted((int(k), *r) for k, lst in selected.items() for r in lst) == sorted(tuple(r) for r in g["selected_rows"].tolist())
    if "image_features" in g.files:
        imf = sel._get_features_for_images(model(), paths)
        np.testing.assert_allclose(imf.cpu().numpy(), g["image_features"], rtol=1e-5, atol=1e-6)
        rep = sel.get_representative_images(model(), paths, paths[1::2])
        assert _idx(rep) == g["representative_images"].tolist()


def test_adversarially_vulnerable_samples_match_reference():
    """accuracy.py:73-96: gradient norm of the error head with respect to its input (network forward + backward in
    PyTorch on the device, TF32 off), masking + image mean + ranking on the scoring path."""
    g = G.load("accuracy_adv")
    seed, N, C, S, block, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, 1, C, S, S, block, g["logits_sha"])
    pool = fakes.Pool(logits, labels)
    unet = torch.nn.Sequential(torch.nn.Conv2d(C + 3, 8, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(8, 2, 3, padding=1))
    unet.load_state_dict({k_[2:].replace("_", ".", 1): torch.from_numpy(g[k_]) for k_ in g.files if k_.startswith("w_")})
    unet = unet.cuda()

    class AdvModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.unet = unet

        @property
        def module(self):
            return self

        def forward(self, x):
            gs = [int(round(float(v) / fakes.GID_SCALE)) for v in x[:, 0, 0, 0].cpu()]
            seg = torch.from_numpy(np.stack([logits[i, 0] for i in gs])).cuda()
            return seg, self.unet(torch.cat([torch.softmax(seg, dim=1), x], dim=1))

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        sel = _factory("accuracy_labels", C, pool, S, bs)
        chosen = sel.get_adversarially_vulnarable_samples(AdvModel(), _paths(N), k)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert _idx(chosen) == g["selected"].tolist()
    np.testing.assert_allclose(sel.last_scores, g["scores"], rtol=2e-5, atol=1e-7)


def test_square_nms_leaves_the_callers_cuda_tensor_like_the_reference():
    """mc_dropout.py:97-103 zeroes the windows of the picks it TAKES, nothing else - also when the caller hands in a
    contiguous CUDA tensor (the NMS kernel itself zeroes the windows of every image-local pick)."""
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCDropout
    from oracle import restate as R
    rng = np.random.default_rng(8)
    m = rng.random((4, 20, 24)).astype(np.float32)
    want_maps = m.copy()
    want_sel, want_count = R.square_nms(want_maps, 5, 6.2)        # 7 picks: fewer than the image-local sequences hold
    for dev in ("cuda", "cpu"):
        t = torch.from_numpy(m.copy()).to(dev)
        got_sel, got_count = ActiveSelectionMCDropout.square_nms(t, 5, 6.2)
        assert (got_sel, got_count) == (want_sel, want_count)
        np.testing.assert_array_equal(t.cpu().numpy(), want_maps)


def test_config1_live_dropout_model_under_dataparallel():
    """BASELINE config 1 as the reference runs it (active_train.py:48,83-85,333; mc_dropout.py:175-178): a stochastic
    nn.Module with live Dropout2d wrapped in nn.DataParallel - no replay.  A forward hook captures the logits of every
    pass as the scorer saw them; the oracle on exactly those logits must give the selector's scores and ranking."""
    N, T, C, S, bs, k = 6, 5, 21, 65, 4, 3

    class SmallNet(torch.nn.Module):            # conv -> Dropout2d -> conv, like the decoder tail (models/decoder.py:35)
        def __init__(self):
            super().__init__()
            torch.manual_seed(3)
            self.body = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 3, padding=1), torch.nn.ReLU(),
                                            torch.nn.Dropout2d(0.25), torch.nn.Conv2d(16, C, 3, padding=1))

        def forward(self, x):
            return self.body(x)

    net = SmallNet().cuda()
    model = torch.nn.DataParallel(net, device_ids=[0]).eval()
    seen = []
    net.register_forward_hook(lambda m, i, o: seen.append(o.detach().clone()))
    g = torch.Generator().manual_seed(9)
    images = torch.rand((N, 3, S, S), generator=g)
    labels = synth.pool_labels(5, list(range(N)), S, S, C, 8)

    class DS(torch.utils.data.Dataset):
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            j = int(self.paths[i])
            return {"image": images[j], "label": torch.from_numpy(labels[j])}

    from deep_active_semantic_segmentation_b200.active_selection import base
    base.paths_dataset.PathsDataset = DS            # (the autouse fixture restores it)
    _set_T(T)
    sel = _factory("variance", C, None, S, bs)
    chosen = sel.get_vote_entropy_for_images(model, _paths(N), k)
    assert not net.body[2].training                                   # model.eval() restored (mc_dropout.py:194)
    assert len(seen) == T * -(-N // bs)
    from oracle import restate as R
    scores, pos = [], 0
    for b0 in range(0, N, bs):
        nb = min(bs, N - b0)
        stack = torch.stack(seen[pos:pos + T], dim=1).cpu().numpy()   # [nb,T,C,S,S]
        pos += T
        assert not np.array_equal(stack[:, 0], stack[:, 1])           # dropout was live: passes differ
        for i in range(nb):
            scores.append(R.image_scores(R.mc_maps(stack[i], labels[b0 + i], C))["vote_entropy"])
    np.testing.assert_allclose(sel.last_scores, scores, rtol=RTOL, atol=ATOL)
    assert _idx(chosen) == R.rank_topk(scores, k, True)
    assert max(scores) > 0


def test_config4_selector_is_ceal_at_one_pass_and_the_oracle_at_many(monkeypatch):
    """ActiveSelectionMCNoise.get_mc_scores_for_images_with_input_noise (BASELINE config 4, composed: mc_noise.py:26-27 +
    ceal.py): with one pass and sigma = 0 it must reproduce the reference's CEAL goldens; with T passes of real input
    noise its scores are the restatement's on exactly the logits the model produced."""
    from deep_active_semantic_segmentation_b200.active_selection import mc_noise
    g = G.load("mc_small")
    seed, N, T, C, H, W, block, k, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, H, W, block, g["logits_sha"])
    pool = fakes.Pool(logits, labels)
    _set_T(1)
    monkeypatch.setattr(mc_noise, "INPUT_NOISE_SIGMA", 0.0)
    sel = _factory("noise_image", C, pool, H, bs)
    for score, key, desc in (("pred_entropy", "ceal_entropy", True), ("confidence", "ceal_conf", False), ("margin", "ceal_margin", False)):
        chosen, allv = sel.get_mc_scores_for_images_with_input_noise(fakes.ReplayModel(pool), _paths(N), k, score=score)
        assert _idx(chosen) == g[key + "_selected"].tolist(), score
        np.testing.assert_allclose(allv[score], g[key], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(sel.last_scores, g[key], rtol=RTOL, atol=ATOL)
    with pytest.raises(NotImplementedError):
        sel.get_mc_scores_for_images_with_input_noise(fakes.ReplayModel(pool), _paths(N), k, score="nope")
    monkeypatch.undo()

    # T passes with the real sigma through a model whose logits depend on its (noisy) input
    Tn, Cn, S = 6, 5, 40
    proj = torch.linspace(-3, 3, Cn * 3).reshape(Cn, 3, 1, 1).cuda()

    class Probe(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.out = []

        def forward(self, x):
            y = torch.nn.functional.conv2d(x, proj)
            self.out.append(y.detach().clone())
            return y

    gen = torch.Generator().manual_seed(4)
    images = torch.rand((5, 3, S, S), generator=gen)
    lab = synth.pool_labels(6, list(range(5)), S, S, Cn, 8)

    class DS(torch.utils.data.Dataset):
        def __init__(self, env, paths, crop_size, include_labels=False):
            self.paths = paths

        def __len__(self):
            return len(self.paths)

        def __getitem__(self, i):
            j = int(self.paths[i])
            return {"image": images[j], "label": torch.from_numpy(lab[j])}

    from deep_active_semantic_segmentation_b200.active_selection import base
    base.paths_dataset.PathsDataset = DS
    _set_T(Tn)
    sel = _factory("noise_image", Cn, None, S, 2)
    model = Probe().cuda().eval()
    chosen, allv = sel.get_mc_scores_for_images_with_input_noise(model, _paths(5), 3, score="margin")
    assert len(model.out) == Tn * 3 and not model.drop.training
    from oracle import restate as R
    want = {kk: [] for kk in R.SCORE_NAMES}
    pos = 0
    for b0 in range(0, 5, 2):
        nb = min(2, 5 - b0)
        stack = torch.stack(model.out[pos:pos + Tn], dim=1).cpu().numpy()
        pos += Tn
        assert not np.array_equal(stack[:, 0], stack[:, 1])            # fresh noise per pass
        for i in range(nb):
            sc = R.image_scores(R.mc_maps(stack[i], lab[b0 + i], Cn))
            for kk in want:
                want[kk].append(sc[kk])
    for kk in ("pred_entropy", "confidence", "margin", "vote_entropy", "expected_entropy"):
        np.testing.assert_allclose(allv[kk], want[kk], rtol=RTOL, atol=1e-6, err_msg=kk)
    assert _idx(chosen) == R.rank_topk(want["margin"], 3, False)        # margin ranks ascending (ceal.py:97)


def test_selection_count_above_the_topk_kernel_limit_falls_back_to_a_stable_host_sort():
    """The reference's sorted()[:k] has no limit; one K3 launch ranks at most 4096 winners.  k = 4500 of 5000 tiny images
    must still be the stable ranking of the pool."""
    from oracle import restate as R
    N, C, S, bs, k = 5000, 3, 4, 500, 4500
    rng = np.random.default_rng(17)
    logits = np.round(rng.standard_normal((N, 1, C, S, S)), 1).astype(np.float32)     # coarse values: many tied scores
    logits[100:200] = logits[300:400]                                                  # and exact duplicates
    labels = np.zeros((N, S, S), dtype=np.float32)
    pool = fakes.Pool(logits, labels)
    _set_T(1)
    sel = _factory("ceal_confidence", C, pool, S, bs)
    chosen = sel.get_least_confident_samples(fakes.ReplayModel(pool), _paths(N), k)
    want_scores = [R.image_scores(R.mc_maps(logits[i], labels[i], C))["confidence"] for i in range(N)]
    np.testing.assert_allclose(sel.last_scores, want_scores, rtol=RTOL, atol=ATOL)
    got = _idx(chosen)
    assert len(got) == k and len(set(got)) == k
    # the mirror ranks ITS float32 scores; ties (duplicates) must come out in index order like Python's stable sort
    assert got == R.rank_topk(list(sel.last_scores), k, False)
    small = sel.get_least_confident_samples(fakes.ReplayModel(pool), _paths(N), 50)
    assert _idx(small) == got[:50]                                                     # device path agrees with the host path
