"""GPU parity of the fused final-upsample scoring kernel (das_mc_upsample_accumulate_finalize):
low-resolution decoder logits in, the scores of `F.interpolate(low_res_x, size, bilinear, align_corners=True)`
(models/deeplab.py:59) followed by the selectors' reductions out."""
import numpy as np
import pytest
import torch

from deep_active_semantic_segmentation_b200 import synth
from deep_active_semantic_segmentation_b200._lib import SCORE_INDEX, DasError
from oracle import restate as R
from tests import golden_util as G

pytestmark = pytest.mark.gpu

RTOL = 1e-5      # north_star tolerance (relative, float32)
ATOL_MAP = 2e-6  # absolute floor for per-pixel maps (as in test_gpu_mc.py)
ATOL_SCORE = 2e-7
PROB_MAPS = ("pred_entropy", "bald", "confidence", "margin")


def _ops():
    from deep_active_semantic_segmentation_b200 import ops
    return ops


def run_up(low, labels, H, W, votes=True, probs=True, weak=False):
    """low numpy [B,T,C,h,w] -> numpy outputs of the fused-upsample CUDA path."""
    ops = _ops()
    B, T, C, h, w = low.shape
    st = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, single_shot=True)
    dev = [torch.from_numpy(np.ascontiguousarray(low[:, t])).cuda() for t in range(T)]
    maps = (["vote_entropy"] if votes else []) + (list(PROB_MAPS) if probs else [])
    lab = None if labels is None else torch.from_numpy(labels).cuda()
    out = st.score_upsampled(dev, lab, maps=maps, scores=True, weak_labels=weak and votes)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def near_tie_mask(up, tol=1e-5):
    """pixels where, in some pass, the two largest upsampled logits are closer than `tol` (relative to their
    size): there a last-bit difference in the interpolation may legitimately flip the vote."""
    part = np.partition(up, up.shape[-3] - 2, axis=-3)
    gap = part[..., -1, :, :] - part[..., -2, :, :]
    scale = np.maximum(np.abs(part[..., -1, :, :]), 1.0)
    return (gap <= tol * scale).any(axis=0)


def check_up_against_oracle(res, low, labels, H, W, votes=True, probs=True):
    B, T, C, h, w = low.shape
    for b in range(B):
        up = R.bilinear_upsample_align_corners(low[b], H, W)
        lab = None if labels is None else labels[b]
        o = R.mc_maps(up, lab, C)
        osc = R.image_scores(o)
        loose = near_tie_mask(up)
        assert loose.mean() < 1e-3
        if votes:
            ok = np.isclose(res["vote_entropy"][b], o["vote_entropy"], rtol=RTOL, atol=ATOL_MAP)
            assert (ok | loose).all(), f"{(~(ok | loose)).sum()} vote-entropy pixels differ away from ties"
            np.testing.assert_allclose(res["scores"][b, SCORE_INDEX["vote_entropy"]], osc["vote_entropy"], rtol=RTOL,
                                       atol=ATOL_SCORE + 2.0 * loose.sum() / loose.size)
            if "weak_labels" in res:
                want = o["votes"][0].copy()
                if lab is not None:
                    want[~R.valid_mask(lab, C)] = 255
                assert ((res["weak_labels"][b] == want) | loose).all()
        else:
            assert np.isnan(res["scores"][b, SCORE_INDEX["vote_entropy"]])
        if probs:
            for name in ("pred_entropy", "confidence", "margin"):
                np.testing.assert_allclose(res[name][b], o[name], rtol=RTOL, atol=ATOL_MAP, err_msg=name)
            np.testing.assert_allclose(res["bald"][b], o["bald"], rtol=RTOL,
                                       atol=ATOL_MAP + RTOL * float(o["pred_entropy"].max()))
            for name in ("pred_entropy", "confidence", "margin", "expected_entropy"):
                np.testing.assert_allclose(res["scores"][b, SCORE_INDEX[name]], osc[name], rtol=RTOL, atol=ATOL_SCORE,
                                           err_msg=name)
        else:
            assert np.isnan(res["scores"][b, SCORE_INDEX["bald"]])


@pytest.mark.parametrize("name", ["upsample_odd", "upsample_rect", "upsample_mid"])
def test_golden_fixtures_through_the_fused_kernel(name):
    """the reference selectors' outputs on F.interpolate(low_res_x) (tests/golden/upsample_*.npz)"""
    g, m, low, labels = G.upsample_case(name)
    res = run_up(low, labels, m["H"], m["W"], weak=True)
    check_up_against_oracle(res, low, labels, m["H"], m["W"])
    nb = g["ve_maps"].shape[0]
    bad = ~np.isclose(res["vote_entropy"][:nb], g["ve_maps"], rtol=RTOL, atol=ATOL_MAP)
    assert bad.sum() <= 2, bad.sum()                      # a tie within one ulp may flip a vote (ATen's small-case path)
    vtol = 1e-5 if name == "upsample_rect" else ATOL_SCORE
    np.testing.assert_allclose(res["scores"][:, SCORE_INDEX["vote_entropy"]], g["ve_scores"], rtol=RTOL, atol=vtol)
    assert R.rank_topk(res["scores"][:, SCORE_INDEX["vote_entropy"]].tolist(), m["k"], True) == g["ve_selected"].tolist()
    # CEAL = the single-pass case
    res1 = run_up(low[:, :1], labels, m["H"], m["W"], votes=False)
    np.testing.assert_allclose(res1["scores"][:, SCORE_INDEX["pred_entropy"]], g["ceal_entropy"], rtol=RTOL, atol=ATOL_SCORE)
    np.testing.assert_allclose(res1["scores"][:, SCORE_INDEX["confidence"]], g["ceal_conf"], rtol=RTOL, atol=ATOL_SCORE)
    np.testing.assert_allclose(res1["scores"][:, SCORE_INDEX["margin"]], g["ceal_margin"], rtol=RTOL, atol=ATOL_SCORE)
    assert R.rank_topk(res1["scores"][:, SCORE_INDEX["margin"]].tolist(), m["k"], False) == g["ceal_margin_selected"].tolist()


CASES = [
    # (B, T, C, h, w, H, W)
    (2, 5, 21, 17, 17, 65, 65),        # odd W: scalar stores, ragged tiles (65 = 4 * 16 + 1)
    (3, 20, 19, 8, 16, 32, 64),        # Cityscapes ratio (in-1)/(out-1) < 1/4, exact tiles
    (1, 1, 19, 5, 5, 16, 16),          # one tile, single pass
    (2, 3, 2, 9, 13, 33, 50),          # two classes, even W that is not a multiple of 16
    (1, 4, 32, 9, 9, 40, 40),          # maximum class count (2 CTAs / SM configuration)
    (2, 6, 11, 6, 7, 41, 47),          # factor ~8 (models/fastscnn.py:22 upsamples its 1/8 classifier output)
    (1, 32, 19, 12, 12, 48, 48),       # maximum passes per launch
    (1, 3, 32, 33, 33, 129, 129),      # wide output -> the 15-warp CTA (16 x 60 tiles), maximum class count
    (2, 2, 2, 9, 61, 33, 241),         # 15-warp CTA with two classes; 241 = 4 * 60 + 1: a tile with one active column
]


@pytest.mark.parametrize("B,T,C,h,w,H,W", CASES)
def test_fused_upsample_matches_oracle(B, T, C, h, w, H, W):
    low = synth.pool_logits(7, list(range(B)), T, C, h, w, 2)
    labels = synth.pool_labels(7, list(range(B)), H, W, C, 8)
    res = run_up(low, labels, H, W, weak=True)
    check_up_against_oracle(res, low, labels, H, W)


@pytest.mark.parametrize("votes,probs", [(True, False), (False, True)])
def test_flag_subsets_and_no_labels(votes, probs):
    B, T, C, h, w, H, W = 2, 5, 19, 9, 17, 36, 68
    low = synth.pool_logits(8, list(range(B)), T, C, h, w, 2)
    res = run_up(low, None, H, W, votes=votes, probs=probs)
    check_up_against_oracle(res, low, None, H, W, votes=votes, probs=probs)


def test_pascal_shape_against_oracle():
    """BASELINE config 1 / 4 shape: 129 x 129 decoder logits -> 513 x 513, C = 21 (planes are not 16-byte aligned)"""
    B, T, C, h, w, H, W = 2, 5, 21, 129, 129, 513, 513
    low = synth.pool_logits(9, list(range(B)), T, C, h, w, 8)
    labels = synth.pool_labels(9, list(range(B)), H, W, C, 32)
    res = run_up(low, labels, H, W, weak=True)
    check_up_against_oracle(res, low, labels, H, W)


def test_full_size_agrees_with_interpolate_then_score():
    """BASELINE config 2 shape (128 x 256 -> 512 x 1024, C = 19, T = 20): the fused kernel against
    F.interpolate on the device followed by the resident-logits kernel (das_mc_accumulate_finalize)."""
    ops = _ops()
    B, T, C, h, w, H, W = 2, 20, 19, 128, 256, 512, 1024
    low, lab = synth.device_pass_logits(11, 0, B, T, C, h, w, "cuda", block=8)
    lab = torch.nn.functional.interpolate(lab[:, None], size=(H, W), mode="nearest")[:, 0].contiguous()
    up = [torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=True) for x in low]
    a = ops.MCState(B, C, H, W, T, single_shot=True).score(up, lab, maps=ops.MAP_NAMES, weak_labels=True)
    b = ops.MCState(B, C, H, W, T, single_shot=True).score_upsampled(low, lab, maps=ops.MAP_NAMES, weak_labels=True)
    torch.cuda.synchronize()
    n = B * H * W
    # ATen's CUDA kernel may round the interpolation differently from its CPU kernel (which we match bit for bit):
    # continuous maps agree to a few ulp of the logits, votes may flip at exact near-ties only
    for name in PROB_MAPS:
        torch.testing.assert_close(b[name], a[name], rtol=1e-4, atol=2e-5, msg=name)
    assert int((a["weak_labels"] != b["weak_labels"]).sum()) <= max(2, n // 100000)
    flips = int((~torch.isclose(a["vote_entropy"], b["vote_entropy"], rtol=1e-5, atol=2e-6)).sum())
    assert flips <= max(4, n // 20000), flips
    torch.testing.assert_close(b["scores"], a["scores"], rtol=1e-4, atol=1e-5)
    # size-independent properties on the fused result
    assert float(b["bald"].min()) > -1e-4
    inval = (lab < 0) | (lab >= C)
    assert float(b["vote_entropy"][inval].abs().max()) == 0.0 and float(b["confidence"][inval].min()) == 1.0
    assert bool((b["weak_labels"][inval] == 255).all())


VARIANT_CASES = [CASES[0], CASES[1], CASES[3], CASES[4], CASES[5], CASES[8], (1, 5, 21, 33, 33, 129, 129),
                 (1, 3, 24, 33, 33, 129, 129)]


@pytest.mark.parametrize("B,T,C,h,w,H,W", VARIANT_CASES)
def test_every_kernel_variant_gives_the_same_maps(B, T, C, h, w, H, W):
    """DAS_OPT_MC_UP_WARPS: 4 / 15 = pixel pairs per lane (csrc/mc_up.cuh), 220 / 216 = one pixel per lane with class
    pairs in the packed fp32 pipe (csrc/mc_up1.cuh).  The library picks one per class count and width; forced one by one
    they all give the oracle's results and - because every sum keeps its association - BIT-identical per-pixel maps (odd
    and even class counts, ragged tiles, two classes, the maximum class count, strips past the right edge of the plane)."""
    ops = _ops()
    low = synth.pool_logits(17, list(range(B)), T, C, h, w, 2)
    labels = synth.pool_labels(17, list(range(B)), H, W, C, 8)
    ref = run_up(low, labels, H, W, weak=True)            # the library's own choice
    check_up_against_oracle(ref, low, labels, H, W)
    tested = 0
    for variant in (4, 15, 220, 216):
        with ops.option("mc_up_warps", variant):
            if not ops.upsample_supported(h, w, H, W):     # e.g. factor 3.85 on 60-column tiles needs an 18th column
                continue
            res = run_up(low, labels, H, W, weak=True)
            tested += 1
        for name in ("vote_entropy", "weak_labels") + PROB_MAPS:
            np.testing.assert_array_equal(res[name], ref[name], err_msg=f"{name}, variant {variant}")
        # the image sums run over another tile partition
        np.testing.assert_allclose(res["scores"], ref["scores"], rtol=2e-6, atol=1e-7, err_msg=f"variant {variant}")
    assert tested >= 3


def test_batching_invariance():
    """scores of an image do not depend on the batch it is scored in (fixed-order tile partials)"""
    B, T, C, h, w, H, W = 3, 4, 19, 17, 33, 65, 130
    low = synth.pool_logits(12, list(range(B)), T, C, h, w, 2)
    labels = synth.pool_labels(12, list(range(B)), H, W, C, 8)
    full = run_up(low, labels, H, W)
    for b in range(B):
        one = run_up(low[b:b + 1], labels[b:b + 1], H, W)
        np.testing.assert_array_equal(one["scores"][0], full["scores"][b])
        np.testing.assert_array_equal(one["pred_entropy"][0], full["pred_entropy"][b])


def test_unsupported_factor_and_bad_arguments_are_loud():
    ops = _ops()
    assert ops.upsample_supported(128, 256, 512, 1024) and ops.upsample_supported(129, 129, 513, 513)
    assert not ops.upsample_supported(32, 32, 64, 64)      # factor 2: a 16-pixel tile needs 9 source rows
    st = ops.MCState(1, 19, 64, 64, 2, single_shot=True)
    with pytest.raises(DasError):
        st.score_upsampled([torch.zeros(1, 19, 32, 32, device="cuda")] * 2, None)
    with pytest.raises(DasError):
        st.score_upsampled([torch.zeros(1, 19, 16, 16)] * 2, None)               # CPU tensor: no CPU path
    with pytest.raises(DasError):
        st.score_upsampled([torch.zeros(1, 18, 16, 16, device="cuda")] * 2, None)  # class count mismatch
    with pytest.raises(DasError):
        st.score_upsampled([torch.zeros(1, 19, 16, 16, device="cuda")] * 3, None)  # more passes than T_cap
