"""GPU parity AT THE BASELINE SIZES (BASELINE.json configs 2, 3, 5) against goldens that do not come from this repo's
CUDA path:

  mc_baseline       2 images of 512 x 1024, C = 19, T = 20 through the REFERENCE selectors (vote entropy scores + sampled
                    map pixels, single-pass CEAL scores); the composed T = 20 scores come from oracle/restate.py
  region_rect       config 3's rectangular 385 x 897 score maps (R = 128) - restatement-pinned: the reference's
                    create_region_maps is square-only
  coreset_baseline  _select_batch of the REFERENCE (sklearn float64) at N = 10 000, D = 2048, L = 50, K = 500

The CUDA side runs exactly what bench.py times: the TMA 3-D-map kernel, create_region_maps, kcenter_greedy with the
tcgen05 filter and the cluster kernel - asserted through the launch counter / filter statistics / handle options."""
import numpy as np
import pytest
import torch

from deep_active_semantic_segmentation_b200 import synth
from tests import fakes
from tests import golden_util as G

pytestmark = pytest.mark.gpu

RTOL = 1e-5            # north_star: per-pixel / per-image scores within 1e-5 relative
ATOL_MAP = 2e-6        # absolute floor of a map value (entropies of confident pixels are ~1e-4)
ATOL_SCORE = 2e-7


@pytest.fixture(autouse=True)
def _synthetic_data_layer():
    from deep_active_semantic_segmentation_b200 import constants
    from deep_active_semantic_segmentation_b200.active_selection import base
    old, old_T = base.paths_dataset.PathsDataset, constants.MC_STEPS
    base.paths_dataset.PathsDataset = fakes.SyntheticPathsDataset
    yield
    base.paths_dataset.PathsDataset, constants.MC_STEPS = old, old_T


def _set_T(T):
    import sys
    from deep_active_semantic_segmentation_b200 import constants
    constants.MC_STEPS = T
    if "constants" in sys.modules and hasattr(sys.modules["constants"], "MC_STEPS"):
        sys.modules["constants"].MC_STEPS = T


@pytest.fixture(scope="module")
def mc_baseline():
    g = G.load("mc_baseline")
    seed, N, T, C, H, W, block, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, H, W, block, g["logits_sha"])
    assert G.sha(labels) == str(g["labels_sha"])
    return g, (seed, N, T, C, H, W, block, bs), logits, labels


def test_tma_3d_kernel_at_config2_size_matches_reference_and_restatement(mc_baseline):
    """das_mc_accumulate_finalize, all T = 20 passes in one launch -> mc_score_tma_kernel<19> (3-D maps)."""
    from deep_active_semantic_segmentation_b200 import _lib, ops
    from deep_active_semantic_segmentation_b200._lib import SCORE_INDEX as S
    g, (seed, N, T, C, H, W, block, bs), logits, labels = mc_baseline
    assert _lib.get_option("mc_tma") == 1
    dev = [torch.from_numpy(np.ascontiguousarray(logits[:, t])).cuda() for t in range(T)]
    lab = torch.from_numpy(labels).cuda()
    n0 = _lib.launch_count()
    st = ops.MCState(N, C, H, W, T, votes=True, probs=True, single_shot=True)
    out = st.score(dev, lab, maps=ops.MAP_NAMES, scores=True)
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 == 2            # the fused TMA kernel + the partial reduction, nothing else
    sc = out["scores"].cpu().numpy()
    rows, cols = g["px_rows"].astype(np.int64), g["px_cols"].astype(np.int64)
    # reference-pinned: vote entropy image scores and map pixels
    np.testing.assert_allclose(sc[:, S["vote_entropy"]], g["ve_scores"], rtol=RTOL, atol=ATOL_SCORE)
    ve = out["vote_entropy"].cpu().numpy()
    np.testing.assert_allclose(ve[:, rows, cols], g["ve_px"], rtol=RTOL, atol=ATOL_MAP)
    assert (g["ve_px"] > 0).sum() > 200             # the sample is not all unanimous pixels
    # restatement-pinned (composed): softmax-mean scores over the T passes
    for name in ("pred_entropy", "confidence", "margin", "expected_entropy"):
        np.testing.assert_allclose(sc[:, S[name]], g["composed_" + name], rtol=RTOL, atol=ATOL_SCORE, err_msg=name)
    np.testing.assert_allclose(sc[:, S["bald"]], g["composed_bald"], rtol=RTOL, atol=RTOL * float(g["composed_pred_entropy"].max()))
    for name in ("pred_entropy", "confidence", "margin"):
        np.testing.assert_allclose(out[name].cpu().numpy()[:, rows, cols], g["composed_px_" + name], rtol=RTOL, atol=ATOL_MAP,
                                   err_msg=name)
    np.testing.assert_allclose(out["bald"].cpu().numpy()[:, rows, cols], g["composed_px_bald"], rtol=RTOL,
                               atol=ATOL_MAP + RTOL * float(g["composed_px_pred_entropy"].max()))
    # the LDG kernel on the same inputs: bit-identical maps (same arithmetic, same partition)
    with ops.option("mc_tma", 0):
        out_ldg = ops.MCState(N, C, H, W, T, votes=True, probs=True, single_shot=True).score(dev, lab, maps=ops.MAP_NAMES)
    for k in ops.MAP_NAMES:
        assert torch.equal(out[k], out_ldg[k]), k


def test_selectors_at_config2_size_match_reference(mc_baseline):
    """The selector mirror end to end on the same pool: MC-dropout ranking and scores, CEAL (T = 1) scores."""
    g, (seed, N, T, C, H, W, block, bs), logits, labels = mc_baseline
    from deep_active_semantic_segmentation_b200.active_selection import get_active_selection_class
    pool = fakes.Pool(logits, labels)
    paths = [str(i) for i in range(N)]
    _set_T(T)
    sel = get_active_selection_class("variance", C, pool, -1, bs)
    chosen = sel.get_vote_entropy_for_images(fakes.ReplayModel(pool), paths, N)
    assert [int(p) for p in chosen] == g["ve_selected"].tolist()
    np.testing.assert_allclose(sel.last_scores, g["ve_scores"], rtol=RTOL, atol=ATOL_SCORE)
    chosen, allv = sel.get_mc_scores_for_images(fakes.ReplayModel(pool), paths, N, score="bald")
    np.testing.assert_allclose(allv["vote_entropy"], g["ve_scores"], rtol=RTOL, atol=ATOL_SCORE)
    np.testing.assert_allclose(allv["pred_entropy"], g["composed_pred_entropy"], rtol=RTOL, atol=ATOL_SCORE)
    assert [int(p) for p in chosen] == list(np.argsort(-g["composed_bald"], kind="stable"))
    ceal = get_active_selection_class("ceal_entropy", C, pool, -1, bs)
    _, ent = ceal.get_maximum_entropy_samples(fakes.ReplayModel(pool), paths, N)
    np.testing.assert_allclose(ent, g["ceal_entropy"], rtol=RTOL, atol=ATOL_SCORE)
    ceal.get_least_confident_samples(fakes.ReplayModel(pool), paths, N)
    np.testing.assert_allclose(ceal.last_scores, g["ceal_conf"], rtol=RTOL, atol=ATOL_SCORE)
    ceal.get_least_margin_samples(fakes.ReplayModel(pool), paths, N)
    np.testing.assert_allclose(ceal.last_scores, g["ceal_margin"], rtol=RTOL, atol=ATOL_SCORE)


def test_create_region_maps_at_config3_rectangular_size():
    """512 x 1024 planes, R = 128 -> 385 x 897 score maps through create_region_maps (and its pieces)."""
    from deep_active_semantic_segmentation_b200 import ops
    from deep_active_semantic_segmentation_b200.active_selection import get_active_selection_class
    from oracle import restate as R
    g = G.load("region_rect")
    seed, N, T, C, H, W, block, Rs, sel_size = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, H, W, block, g["logits_sha"])
    existing = G.regions_from_rows(g["existing"], N)
    pool = fakes.Pool(logits, labels)
    _set_T(T)
    sel = get_active_selection_class("variance", C, pool, -1, 2)
    regions, count = sel.create_region_maps(fakes.ReplayModel(pool), [str(i) for i in range(N)], existing, Rs, sel_size)
    want = G.regions_from_rows(g["regions"], N)
    assert count == int(g["count"]) and count < float(g["K"])          # the 0.01 stop rule fired before K picks
    assert {int(k): v for k, v in regions.items()} == {i: want[i] for i in range(N) if want[i]}
    # pieces: vote entropy -> suppression -> fp64 sliding sums -> pool min-max, on sampled positions
    maps = torch.stack([torch.from_numpy(R.vote_entropy_map(R.votes_from_logits(logits[i]), C, R.valid_mask(labels[i], C)))
                        for i in range(N)]).cuda()
    ops.suppress_rects(maps, [(i, *rc) for i in range(N) for rc in existing[i]])
    mm = ops.new_minmax("cuda")
    sm = ops.box_sum(maps, Rs, mm)
    assert tuple(sm.shape) == (N, H - Rs + 1, W - Rs + 1) == (N, 385, 897)
    np.testing.assert_allclose(mm.cpu().numpy(), [float(g["raw_min"]), float(g["raw_max"])], rtol=1.2e-7)
    ops.minmax_normalise(sm, mm)
    i_, r_, c_ = (g["px"][j].astype(np.int64) for j in range(3))
    np.testing.assert_allclose(sm.cpu().numpy()[i_, r_, c_], g["norm_px"], rtol=RTOL, atol=1e-7)


def test_kcenter_at_config5_size_matches_reference_sklearn_loop():
    """N = 10 000, D = 2048, L = 50, K = 500: tcgen05 distance filter + cluster-resident greedy loop vs the
    reference's sklearn float64 _select_batch (500 picks, bit-exact)."""
    from deep_active_semantic_segmentation_b200 import _lib
    from deep_active_semantic_segmentation_b200.active_selection import get_active_selection_class
    g = G.load("coreset_baseline")
    seed, N, D, L, K = (int(v) for v in g["meta"])
    feats = synth.coreset_features(seed, N, D)
    assert G.sha(feats) == str(g["features_sha"])
    assert _lib.get_option("kc_cluster") == 1 and _lib.get_option("gemm_2cta") == 1
    sel = get_active_selection_class("coreset", 19, None, 513, 4)
    n0 = _lib.launch_count()
    picks = sel._select_batch(feats, list(range(L)), K)
    launches = _lib.launch_count() - n0
    assert picks == g["picks"].tolist()
    assert sel.last_filter_stats is not None and sel.last_filter_stats[0] < 0.05 * sel.last_filter_stats[1]   # filter on
    assert launches <= 6                       # prepare, GEMM, init, ONE cluster kernel (+ stats copy) - not 500 steps
    md = sel.last_min_distances.cpu().numpy()
    np.testing.assert_allclose(md.max(), float(g["min_dist_max"]), rtol=1e-9)
    np.testing.assert_allclose(md[:256], g["min_dist_head"], rtol=1e-9, atol=1e-4)   # sklearn's expanded form: ~1e-6 of noise at d = 0
    # the reference's helper with its own contract (core_set.py:32-38)
    d0 = sel._updated_distances(list(range(L)), feats.astype(np.float64), None)
    assert d0.shape == (N, 1) and d0.dtype == np.float64
    d1 = sel._updated_distances([int(picks[0])], feats.astype(np.float64), d0)
    assert int(np.argmax(d0)) == picks[0] and int(np.argmax(d1)) == picks[1]
