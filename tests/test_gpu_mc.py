"""GPU parity of K1 (accumulate) + K2 (finalize) through the C ABI, against the numpy oracle."""
import numpy as np
import pytest
import torch

from deep_active_semantic_segmentation_b200 import synth
from oracle import restate as R

pytestmark = pytest.mark.gpu

RTOL = 1e-5     # north_star: per-pixel and per-image scores within 1e-5 relative (float32)
ATOL_MAP = 2e-6  # absolute floor for per-pixel maps: entropies of confident pixels are ~1e-4 and the
                 # reference's own float32 formula carries ~1e-7 of rounding there
ATOL_SCORE = 2e-7


def _ops():
    from deep_active_semantic_segmentation_b200 import ops
    return ops


def run_gpu(logits, labels, group, votes=True, probs=True, weak=False, fused=False):
    """logits numpy [B,T,C,H,W] -> dict of numpy outputs from the CUDA path.

    fused=False: das_mc_accumulate per group, then das_mc_finalize (K1 ... K1, K2).
    fused=True : das_mc_accumulate for all but the last group, das_mc_accumulate_finalize for the last one
                 (single-shot state, i.e. no accumulators in HBM at all, when one group holds every pass)."""
    ops = _ops()
    B, T, C, H, W = logits.shape
    single = fused and group >= T and T <= 32
    st = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, single_shot=single)
    dev = [torch.from_numpy(np.ascontiguousarray(logits[:, t])).cuda() for t in range(T)]
    maps = (["vote_entropy"] if votes else []) + ([m for m in ops.MAP_NAMES if m != "vote_entropy"] if probs else [])
    lab = None if labels is None else torch.from_numpy(labels).cuda()
    groups = [dev[t0:t0 + group] for t0 in range(0, T, group)]
    if fused:
        for g in groups[:-1]:
            st.accumulate(g)
        out = st.score(groups[-1], lab, maps=maps, scores=True, weak_labels=weak and votes)
    else:
        for g in groups:
            st.accumulate(g)
        out = st.finalize(lab, maps=maps, scores=True, weak_labels=weak and votes)
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if votes and not single:
        res["votes"] = st.votes_tensor().cpu().numpy()
    torch.cuda.synchronize()
    return res


def check_against_oracle(res, logits, labels, votes=True, probs=True):
    B, T, C, H, W = logits.shape
    from deep_active_semantic_segmentation_b200._lib import SCORE_INDEX
    for b in range(B):
        o = R.mc_maps(logits[b], None if labels is None else labels[b], C)
        osc = R.image_scores(o)
        if votes:
            if "votes" in res:
                np.testing.assert_array_equal(res["votes"][b, :T], o["votes"])        # index work: bit exact
            np.testing.assert_allclose(res["vote_entropy"][b], o["vote_entropy"], rtol=RTOL, atol=ATOL_MAP)
            np.testing.assert_allclose(res["scores"][b, SCORE_INDEX["vote_entropy"]], osc["vote_entropy"], rtol=RTOL, atol=ATOL_SCORE)
        else:
            assert np.isnan(res["scores"][b, SCORE_INDEX["vote_entropy"]])
        if probs:
            for name in ("pred_entropy", "confidence", "margin"):
                np.testing.assert_allclose(res[name][b], o[name], rtol=RTOL, atol=ATOL_MAP, err_msg=name)
            # BALD is a difference of two entropies: tolerance relative to the entropies' scale
            np.testing.assert_allclose(res["bald"][b], o["bald"], rtol=RTOL, atol=ATOL_MAP + RTOL * float(o["pred_entropy"].max()))
            for name in ("pred_entropy", "confidence", "margin", "expected_entropy"):
                np.testing.assert_allclose(res["scores"][b, SCORE_INDEX[name]], osc[name], rtol=RTOL, atol=ATOL_SCORE, err_msg=name)
            np.testing.assert_allclose(res["scores"][b, SCORE_INDEX["bald"]], osc["bald"], rtol=RTOL, atol=RTOL * float(osc["pred_entropy"]))
        else:
            assert np.isnan(res["scores"][b, SCORE_INDEX["bald"]])


CASES = [
    # (B, T, C, H, W, block)   odd H*W -> scalar path, H*W % 4 == 0 -> 128-bit path
    (2, 5, 21, 65, 65, 8),
    (3, 20, 19, 32, 64, 8),
    (1, 1, 19, 16, 16, 4),      # CEAL: single pass
    (2, 7, 2, 9, 7, 3),         # smallest class count, ragged
    (1, 4, 32, 24, 20, 4),      # largest class count
    (2, 33, 5, 12, 12, 4),      # more passes than one launch group can take
]


@pytest.mark.parametrize("B,T,C,H,W,block", CASES)
@pytest.mark.parametrize("group", [1, 3, 64])
@pytest.mark.parametrize("fused", [False, True])
def test_mc_matches_oracle(B, T, C, H, W, block, group, fused):
    gs = list(range(B))
    logits = synth.pool_logits(7, gs, T, C, H, W, block)
    labels = synth.pool_labels(7, gs, H, W, C, block)
    res = run_gpu(logits, labels, min(group, T), fused=fused)
    check_against_oracle(res, logits, labels)


def test_mc_votes_only_and_probs_only_and_no_labels():
    logits = synth.pool_logits(9, [0, 1], 6, 19, 20, 28, 4)
    labels = synth.pool_labels(9, [0, 1], 20, 28, 19, 4)
    for fused in (False, True):
        check_against_oracle(run_gpu(logits, labels, 2, votes=True, probs=False, fused=fused), logits, labels, True, False)
        check_against_oracle(run_gpu(logits, labels, 2, votes=False, probs=True, fused=fused), logits, labels, False, True)
        check_against_oracle(run_gpu(logits, None, 6, fused=fused), logits, None)
        check_against_oracle(run_gpu(logits, labels, 6, votes=True, probs=False, fused=fused), logits, labels, True, False)


def test_mc_exact_ties_pick_first_class_and_labels_edge_values():
    B, T, C, H, W = 1, 3, 7, 8, 8
    logits = np.zeros((B, T, C, H, W), dtype=np.float32)       # every class ties -> vote 0
    logits[0, 1, 3] = 2.0
    logits[0, 1, 5] = 2.0                                       # tie between 3 and 5 -> 3
    labels = np.zeros((B, H, W), dtype=np.float32)
    labels[0, 0, :] = [-1, -0.0, 6, 6.5, 7, 255, np.nan, 3]     # valid iff not (l<0 or l>=C); NaN stays valid
    wl = R.votes_from_logits(logits[0, 0]).copy()
    wl[~R.valid_mask(labels[0], C)] = 255
    for fused, group in ((False, 3), (True, 3), (True, 2)):
        res = run_gpu(logits, labels, group, weak=True, fused=fused)
        check_against_oracle(res, logits, labels)
        if "votes" in res:
            assert (res["votes"][0, 0] == 0).all() and (res["votes"][0, 1] == 3).all()
        np.testing.assert_array_equal(res["weak_labels"][0], wl)


def test_mc_extreme_logits():
    # large magnitudes / wide dynamic range: softmax must stay max-subtracted like ATen's
    rng = np.random.default_rng(3)
    logits = (rng.standard_normal((1, 4, 19, 8, 16)) * 30).astype(np.float32)
    logits[0, 0, 2] += 500.0
    logits[0, 1, 4] -= 500.0
    res = run_gpu(logits, None, 4)
    check_against_oracle(res, logits, None)
    assert np.isfinite(res["scores"]).all()


def test_full_size_properties():
    """BASELINE config-2 shape (512x1024, C=19, T=20): size-independent properties instead of the oracle."""
    ops = _ops()
    from deep_active_semantic_segmentation_b200._lib import SCORE_INDEX as S
    B, T, C, H, W = 2, 20, 19, 512, 1024
    passes, lab = synth.device_pass_logits(1, 0, B, T, C, H, W, "cuda")

    def run(order, group):
        st = ops.MCState(B, C, H, W, T)
        for t0 in range(0, T, group):
            st.accumulate([passes[t] for t in order[t0:t0 + group]])
        out = st.finalize(lab, maps=ops.MAP_NAMES, scores=True)
        return out, st.votes_tensor()

    out, votes = run(list(range(T)), 1)
    # votes are exactly the per-pass argmax (integer work: bit exact, here against torch on the same device)
    for t in (0, 7, 19):
        assert torch.equal(votes[:, t].long(), passes[t].argmax(dim=1))
    # grouping passes differently does not change anything (same per-pixel accumulation order)
    out_g, votes_g = run(list(range(T)), 20)
    assert torch.equal(votes, votes_g)
    for k in list(ops.MAP_NAMES) + ["scores"]:
        assert torch.equal(out[k], out_g[k]), k
    # the fused last-group kernel (single shot, and after a streamed head) gives bit-identical results
    for group in (20, 8):
        st = ops.MCState(B, C, H, W, T, single_shot=(group == 20))
        for t0 in range(0, T - group, group):
            st.accumulate(passes[t0:t0 + group])
        out_f = st.score(passes[T - (T % group or group):], lab, maps=ops.MAP_NAMES, scores=True)
        for k in list(ops.MAP_NAMES) + ["scores"]:
            assert torch.equal(out[k], out_f[k]), (group, k)
    # vote entropy is invariant under a permutation of the passes (histogram), bit for bit
    perm = list(np.random.default_rng(0).permutation(T))
    out_p, _ = run(perm, 4)
    assert torch.equal(out["vote_entropy"], out_p["vote_entropy"])
    torch.testing.assert_close(out["pred_entropy"], out_p["pred_entropy"], rtol=1e-5, atol=2e-6)
    # ranges
    ve, pe, bald, conf, marg = (out[k] for k in ops.MAP_NAMES)
    assert ve.min() >= 0 and ve.max() <= np.log2(min(C, T)) + 1e-5
    assert pe.min() >= 0 and pe.max() <= np.log2(C) + 1e-5
    assert bald.min() >= -1e-5                       # mutual information is non-negative
    assert conf.min() >= 1.0 / C - 1e-6 and conf.max() <= 1 + 1e-6 and marg.min() >= 0 and (marg <= conf + 1e-6).all()
    invalid = (lab < 0) | (lab >= C)
    assert (ve[invalid] == 0).all() and (pe[invalid] == 0).all() and (bald[invalid] == 0).all()
    assert (conf[invalid] == 1).all() and (marg[invalid] == 1).all()
    # image score == mean of the map over all pixels
    for name in ops.MAP_NAMES:
        ref = out[name].double().mean(dim=(1, 2)).float()
        torch.testing.assert_close(out["scores"][:, S[name]], ref, rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(out["scores"][:, S["bald"]],
                               out["scores"][:, S["pred_entropy"]] - out["scores"][:, S["expected_entropy"]], rtol=1e-5, atol=1e-6)


def test_errors_are_loud():
    ops = _ops()
    from deep_active_semantic_segmentation_b200._lib import DasError
    with pytest.raises(DasError):
        ops.MCState(1, 33, 8, 8, 4)               # > DAS_MAX_CLASSES
    with pytest.raises(DasError):
        ops.MCState(1, 5, 8, 8, 256)              # > DAS_MAX_PASSES
    st = ops.MCState(1, 5, 8, 8, 2)
    with pytest.raises(DasError):
        st.accumulate(torch.zeros(1, 5, 8, 8))     # CPU tensor: no CPU path
    with pytest.raises(DasError):
        st.accumulate(torch.zeros(1, 4, 8, 8, device="cuda"))   # wrong class count
    with pytest.raises(DasError):
        st.finalize(None)                          # nothing accumulated
    st.accumulate([torch.zeros(1, 5, 8, 8, device="cuda")] * 2)
    with pytest.raises(DasError):
        st.accumulate(torch.zeros(1, 5, 8, 8, device="cuda"))   # state is full (T_cap = 2)
    single = ops.MCState(1, 5, 8, 8, 2, single_shot=True)
    with pytest.raises(DasError):
        single.accumulate(torch.zeros(1, 5, 8, 8, device="cuda"))   # a single-shot state has no accumulators
    with pytest.raises(DasError):
        ops.MCState(1, 5, 8, 8, 33, single_shot=True)               # more passes than one launch can take


@pytest.mark.parametrize("B,T,C,H,W", [(2, 6, 19, 32, 64), (3, 5, 21, 65, 65), (2, 4, 7, 33, 34), (1, 20, 19, 128, 256)])
@pytest.mark.parametrize("probs,votes", [(True, True), (True, False)])
def test_tma_ring_kernel_is_bit_identical_to_ldg_kernel(B, T, C, H, W, probs, votes, monkeypatch):
    """The TMA-staged single-shot kernel (3-D maps for H*W % 4 == 0, flat shifted 1-D maps otherwise) and the LDG
    kernel run the same per-pixel arithmetic: maps and votes must agree bit for bit, image scores to fp32 rounding
    of the (differently partitioned) block partials."""
    ops = _ops()
    gs = list(range(B))
    logits = synth.pool_logits(21, gs, T, C, H, W, 8)
    labels = synth.pool_labels(21, gs, H, W, C, 8)
    dev = [torch.from_numpy(np.ascontiguousarray(logits[:, t])).cuda() for t in range(T)]
    lab = torch.from_numpy(labels).cuda()
    maps = (["vote_entropy"] if votes else []) + [m for m in ops.MAP_NAMES if m != "vote_entropy"]
    outs = {}
    for tma in ("1", "0"):
        with ops.option("mc_tma", int(tma)):
            n0 = _lib_launches()
            st = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, single_shot=True)
            outs[tma] = st.score(dev, lab, maps=maps, scores=True, weak_labels=votes)
            torch.cuda.synchronize()
            assert _lib_launches() - n0 == 2
    for k in maps + (["weak_labels"] if votes else []):
        assert torch.equal(outs["1"][k], outs["0"][k]), k
    a, b = outs["1"]["scores"].cpu().numpy(), outs["0"]["scores"].cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-7, equal_nan=True)
    for bi in range(B):            # and both agree with the oracle
        o = R.mc_maps(logits[bi], labels[bi], C)
        np.testing.assert_allclose(outs["1"]["pred_entropy"][bi].cpu().numpy(), o["pred_entropy"], rtol=1e-5, atol=2e-6)


def _lib_launches():
    from deep_active_semantic_segmentation_b200 import _lib
    return _lib.launch_count()


@pytest.mark.parametrize("H,W,C,T", [(512, 1024, 19, 20), (513, 513, 21, 20)])
def test_tma_vs_ldg_bitwise_at_baseline_size(H, W, C, T, monkeypatch):
    """BASELINE shapes (Cityscapes 512x1024 C=19, Pascal 513x513 C=21; T=20), inputs generated on the device:
    the TMA ring kernel (3-D maps / flat shifted 1-D maps) and the LDG kernel must produce bit-identical maps."""
    ops = _ops()
    B = 2
    passes, labels = synth.device_pass_logits(77, 0, B, T, C, H, W, torch.device("cuda", 0))
    maps = list(ops.MAP_NAMES)
    outs = {}
    for tma in ("1", "0"):
        with ops.option("mc_tma", int(tma)):
            st = ops.MCState(B, C, H, W, T, votes=True, probs=True, single_shot=True)
            outs[tma] = st.score(passes, labels, maps=maps, scores=True, weak_labels=True)
            torch.cuda.synchronize()
    for k in maps + ["weak_labels"]:
        assert torch.equal(outs["1"][k], outs["0"][k]), k
    np.testing.assert_allclose(outs["1"]["scores"].cpu().numpy(), outs["0"]["scores"].cpu().numpy(), rtol=2e-6, atol=1e-7)
    # sanity of the content: masked border is exactly 0 / 1, BALD >= -eps, unanimous pixels have zero vote entropy
    ve, bald, conf = outs["1"]["vote_entropy"], outs["1"]["bald"], outs["1"]["confidence"]
    assert float(ve[:, :16].abs().max()) == 0.0 and float(conf[:, :16].min()) == 1.0
    assert float(bald.min()) > -1e-5 and float(ve.max()) <= np.log2(min(C, T)) + 1e-5


@pytest.mark.parametrize("C", [2, 3, 5, 8, 13, 16, 20, 22, 24, 25, 28, 31, 32])
def test_every_class_count_family_tma_ldg_oracle(C, monkeypatch):
    """One representative per register / shared-memory configuration of the kernels (C <= 20, 21-24, 25-32; LDG
    accumulators in registers vs shared memory): TMA == LDG bit for bit, both == oracle, on an aligned and an odd plane."""
    ops = _ops()
    for (H, W) in ((16, 32), (23, 29)):
        B, T = 2, 4
        logits = synth.pool_logits(100 + C, [0, 1], T, C, H, W, 8)
        labels = synth.pool_labels(100 + C, [0, 1], H, W, C, 8)
        dev = [torch.from_numpy(np.ascontiguousarray(logits[:, t])).cuda() for t in range(T)]
        lab = torch.from_numpy(labels).cuda()
        outs = {}
        for tma in ("1", "0"):
            with ops.option("mc_tma", int(tma)):
                st = ops.MCState(B, C, H, W, T, votes=True, probs=True, single_shot=True)
                outs[tma] = st.score(dev, lab, maps=list(ops.MAP_NAMES), scores=True)
                torch.cuda.synchronize()
        for k in ops.MAP_NAMES:
            assert torch.equal(outs["1"][k], outs["0"][k]), (C, H, W, k)
        check_against_oracle({k: v.cpu().numpy() for k, v in outs["1"].items()}, logits, labels)


@pytest.mark.parametrize("H,W", [(8, 8), (5, 7), (16, 16), (16, 17)])
@pytest.mark.parametrize("votes,probs", [(True, True), (True, False)])
def test_signed_zero_ties_vote_like_torch_argmax(H, W, votes, probs):
    """-0.0 == +0.0 for torch.argmax (mc_dropout.py:40): the first of them wins.  The kernels take the sign bit of
    x_c + (0 - max) as "not a maximum", so a -0.0 logit next to a +0.0 maximum must still produce +0.
    Shapes cover VEC = 4 / 2 / 1 LDG kernels, the TMA 3-D-map kernel (16 x 16) and the flat TMA kernel (16 x 17)."""
    C, T = 7, 5
    x = np.full((1, T, C, H, W), -3.0, dtype=np.float32)
    nz, pz = np.float32(-0.0), np.float32(0.0)
    x[0, 0, 0], x[0, 0, 1] = nz, pz                   # -0 before +0           -> 0
    x[0, 1, 1], x[0, 1, 2] = pz, nz                   # +0 before -0           -> 1
    x[0, 2, :] = nz                                   # every class -0         -> 0
    x[0, 3, 3], x[0, 3, 5] = nz, pz                   # -0 at 3, +0 at 5       -> 3
    x[0, 4, 2], x[0, 4, 4], x[0, 4, 6] = nz, nz, pz   # two -0 then +0         -> 2
    x[0, 4, 0] = np.float32(-1e-30)                   # a tiny negative is NOT a maximum
    want = torch.argmax(torch.from_numpy(x[0]), dim=1).numpy().astype(np.uint8)
    assert want[:, 0, 0].tolist() == [0, 1, 0, 3, 2]
    np.testing.assert_array_equal(R.votes_from_logits(x[0]), want)
    labels = np.zeros((1, H, W), dtype=np.float32)
    for fused, group in ((False, 1), (False, T), (True, T)):
        res = run_gpu(x, labels, group, votes=votes, probs=probs, weak=True, fused=fused)
        if "votes" in res:
            np.testing.assert_array_equal(res["votes"][0, :T], want)
        np.testing.assert_array_equal(res["weak_labels"][0], want[0])
        check_against_oracle(res, x, labels, votes, probs)


def _truth64(logits, C):
    """float64 evaluation of the composed formulas for one image: [T,C,H,W] -> dict of [H,W]."""
    x = logits.astype(np.float64)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    pb = p.mean(axis=0)
    ent = lambda q: -(q * np.log2(q + 1e-12)).sum(axis=-3)
    part = np.partition(pb, C - 2, axis=0)
    pe, ee = ent(pb), ent(p).mean(axis=0)
    return {"pred_entropy": pe, "expected_entropy": ee, "bald": pe - ee, "confidence": pb.max(axis=0),
            "margin": part[-1] - part[-2]}


def _sweep_cases():
    rng = np.random.default_rng(11)
    H, W = 32, 64
    n = H * W
    cases = {}
    # (a) the largest mean probability sweeps through 0.5, where MUFU.LG2 changes from a relative to an absolute bound
    C, T = 19, 4
    x = np.full((T, C, n), -40.0, dtype=np.float32)
    x[:, 0] = 0.0
    x[:, 1] = np.linspace(-0.5, 0.5, n, dtype=np.float32)[None] + rng.normal(0, 1e-3, (T, n)).astype(np.float32)
    cases["top_prob_through_0.5"] = x.reshape(T, C, H, W)
    # (b) the whole range of a dominant probability: 1e-6 .. 1 - 1e-6
    x = np.full((T, C, n), -25.0, dtype=np.float32)
    x[:, 3] = 0.0
    x[:, 7] = np.linspace(-14.0, 14.0, n, dtype=np.float32)[None]
    cases["top_prob_full_range"] = x.reshape(T, C, H, W)
    # (c) C = 32 near-uniform: every probability close to 1/32, entropies close to 5 bits, margins close to 0
    x = (rng.normal(0, 1e-3, (3, 32, n))).astype(np.float32) + np.float32(7.5)
    cases["c32_near_uniform"] = x.reshape(3, 32, H, W)
    # (d) T = 255 accumulations of wide-range logits (the streaming state in HBM)
    base = rng.normal(0, 4.0, (17, 8, n)).astype(np.float32)
    cases["t255_accumulation"] = base[np.arange(255) % 17].reshape(255, 8, H, W)
    # (e) huge magnitudes: exp2 underflow of everything but the maximum
    x = (rng.normal(0, 200.0, (T, C, n))).astype(np.float32)
    cases["huge_magnitudes"] = x.reshape(T, C, H, W)
    return cases


@pytest.mark.parametrize("case", ["top_prob_through_0.5", "top_prob_full_range", "c32_near_uniform", "t255_accumulation",
                                  "huge_magnitudes"])
def test_approximation_error_sweep(case, capsys):
    """ex2.approx / rcp.approx / lg2.approx (csrc/das_common.cuh) at their worst cases: max error of every softmax-derived
    map against a float64 evaluation, for the streaming, the fused LDG and the TMA kernels.  north_star: 1e-5 relative;
    the absolute floor is a few float32 ulp of the map's scale (entropies of C classes are <= log2 C)."""
    x = _sweep_cases()[case]
    T, C, H, W = x.shape
    truth = _truth64(x, C)
    oracle = R.mc_maps(x, None, C)
    runs = {"streaming G=1" if T <= 32 else "streaming G=32": run_gpu(x[None], None, 1 if T <= 32 else 32, fused=False)}
    if T <= 32:
        runs["fused single shot (TMA)"] = run_gpu(x[None], None, T, fused=True)
        ops = _ops()
        with ops.option("mc_tma", 0):
            runs["fused single shot (LDG)"] = run_gpu(x[None], None, T, fused=True)
    scale = {"pred_entropy": np.log2(C), "expected_entropy": np.log2(C), "bald": np.log2(C), "confidence": 1.0, "margin": 1.0}
    report = []
    for kname, res in runs.items():
        for name in ("pred_entropy", "bald", "confidence", "margin"):
            got = res[name][0].astype(np.float64)
            err = np.abs(got - truth[name])
            rel = err / np.maximum(np.abs(truth[name]), 1e-2 * scale[name])
            o_err = np.abs(oracle[name].astype(np.float64) - truth[name])
            report.append(f"{case:>22} {kname:>24} {name:>13}: max abs {err.max():.2e} max rel(>1% of scale) {rel.max():.2e} "
                          f"(float32 restatement: {o_err.max():.2e})")
            # the bar, written out: 1e-5 relative + 4 float32 ulp of the map scale (BALD: a difference of two such maps)
            atol = (8 if name == "bald" else 4) * 1.2e-7 * scale[name]
            assert (err <= 1e-5 * np.abs(truth[name]) + atol).all(), report[-1]
    with capsys.disabled():
        print("\n" + "\n".join(report))


def test_mc_noise_input_perturbation_is_observed():
    """mc_noise.py:26-27: every pass sees image + N(0, 0.125) with FRESH noise.  A model whose logits are a fixed 1 x 1
    projection of its input records what it was fed: sigma, mean, per-pass freshness, and the scorer's output must be the
    oracle's on exactly those logits."""
    from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionMCNoise, base
    from deep_active_semantic_segmentation_b200 import constants
    C, T, B, H, W = 6, 7, 2, 24, 40
    proj = torch.linspace(-2, 2, C * 3).reshape(C, 3, 1, 1).cuda()

    class Probe(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.drop = torch.nn.Dropout2d(0.25)
            self.seen, self.out = [], []

        def forward(self, x):
            self.seen.append(x.detach().clone())
            y = torch.nn.functional.conv2d(x, proj)
            self.out.append(y.detach().clone())
            return y

    g = torch.Generator(device="cuda").manual_seed(3)
    images = torch.rand((B, 3, H, W), generator=g, device="cuda")
    labels = torch.zeros((B, H, W), device="cuda")
    old_T = constants.MC_STEPS
    constants.MC_STEPS = T
    import sys
    ref_constants = sys.modules.get("constants")
    old_ref = getattr(ref_constants, "MC_STEPS", None) if ref_constants is not None else None
    if old_ref is not None:
        ref_constants.MC_STEPS = T
    try:
        sel = ActiveSelectionMCNoise(C, None, -1, B)
        model = Probe().cuda().eval()
        maps = sel._get_vote_entropy_for_batch_with_input_noise(model, images, labels)
    finally:
        constants.MC_STEPS = old_T
        if old_ref is not None:
            ref_constants.MC_STEPS = old_ref
    assert len(model.seen) == T and len(maps) == B
    noise = torch.stack([s - images for s in model.seen])              # [T,B,3,H,W]
    n = noise[0].numel()
    assert abs(float(noise.std()) - 0.125) < 0.125 * 0.02              # sigma of mc_noise.py:26 (T*n = 40k samples)
    assert abs(float(noise.mean())) < 4 * 0.125 / np.sqrt(T * n)
    for t in range(T):                                                # every pass is perturbed, and differently
        assert abs(float(noise[t].std()) - 0.125) < 0.125 * 0.06
        for u in range(t):
            c = float((noise[t] * noise[u]).mean()) / 0.125 ** 2
            assert abs(c) < 6 / np.sqrt(n), (t, u, c)
    logits = torch.stack(model.out, dim=1).cpu().numpy()               # [B,T,C,H,W] as the scorer saw them
    for b in range(B):
        o = R.mc_maps(logits[b], labels[b].cpu().numpy(), C)
        np.testing.assert_allclose(maps[b].cpu().numpy(), o["vote_entropy"], rtol=RTOL, atol=ATOL_MAP)
    assert float(torch.stack(maps).max()) > 0                          # the noise does flip votes on this model


def test_handle_options_and_info():
    """das_handle: per-device, options readable / writable, L2 geometry and SM count come from the device."""
    import ctypes
    from deep_active_semantic_segmentation_b200 import _lib
    lib = _lib.load()
    h = _lib.handle(0)
    assert _lib.handle(0).value == h.value == _lib.handle(torch.device("cuda", 0)).value          # one handle per device
    assert lib.das_handle_device(h) == 0
    assert lib.das_handle_sm_count(h) == torch.cuda.get_device_properties(0).multi_processor_count
    info = _lib.l2_info(0)
    assert info["l2_bytes"] == torch.cuda.get_device_properties(0).L2_cache_size and info["persisting_max_bytes"] <= info["l2_bytes"]
    old = _lib.set_option("mc_tma_ctas", 2)
    assert _lib.get_option("mc_tma_ctas") == 2 and _lib.set_option("mc_tma_ctas", old) == 2
    assert lib.das_handle_set_option(h, _lib.OPTIONS["mc_up_warps"], 7) == -1                       # only 0 / 4 / 15 / 220 / 216
    assert lib.das_handle_set_option(h, 99, 1) == -1
    v = ctypes.c_int()
    assert lib.das_handle_get_option(h, _lib.OPTIONS["mc_tma"], ctypes.byref(v)) == 0 and v.value in (0, 1)
    # a second, independent handle on the same device keeps its own options; destroying it leaves the first intact
    h2 = ctypes.c_void_p()
    assert lib.das_handle_create(0, ctypes.byref(h2)) == 0 and h2.value != h.value
    assert lib.das_handle_set_option(h2, _lib.OPTIONS["mc_tma"], 0) == 0 and _lib.get_option("mc_tma") == 1
    assert lib.das_handle_destroy(h2) == 0
    assert lib.das_handle_create(torch.cuda.device_count() + 3, ctypes.byref(h2)) == -1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """One handle per device: tensors on cuda:1 are scored while cuda:0 is the current device (the entry points switch
    to the handle's device for the call), with the TMA path (per-handle descriptor cache) and k-center (per-handle
    scratch), and both devices give the oracle's results."""
    ops = _ops()
    from deep_active_semantic_segmentation_b200 import _lib
    B, T, C, H, W = 2, 5, 19, 32, 64
    logits = synth.pool_logits(31, [0, 1], T, C, H, W, 8)
    labels = synth.pool_labels(31, [0, 1], H, W, C, 8)
    torch.cuda.set_device(0)
    assert _lib.handle(0).value != _lib.handle(1).value
    outs = []
    for d in (1, 0, 1):
        dev = torch.device("cuda", d)
        ps = [torch.from_numpy(np.ascontiguousarray(logits[:, t])).to(dev) for t in range(T)]
        st = ops.MCState(B, C, H, W, T, device=dev, single_shot=True)
        out = st.score(ps, torch.from_numpy(labels).to(dev), maps=ops.MAP_NAMES, scores=True)
        assert out["scores"].device == dev and torch.cuda.current_device() == 0
        outs.append({k: v.cpu().numpy() for k, v in out.items()})
        st2 = ops.MCState(B, C, H, W, T, device=dev)            # streaming form on the same device
        for p in ps:
            st2.accumulate(p)
        fin = st2.finalize(torch.from_numpy(labels).to(dev), maps=("pred_entropy",))
        np.testing.assert_array_equal(fin["pred_entropy"].cpu().numpy(), outs[-1]["pred_entropy"])
    for k in outs[0]:
        np.testing.assert_array_equal(outs[0][k], outs[1][k])
        np.testing.assert_array_equal(outs[0][k], outs[2][k])
    check_against_oracle(outs[0], logits, labels)
    feats = synth.coreset_features(2, 600, 64)
    want, _ = R.kcenter_greedy(feats, [0, 1, 2], 12)
    for d in (1, 0):
        f = torch.from_numpy(feats).to(f"cuda:{d}")
        cen = torch.tensor([0, 1, 2], dtype=torch.int32, device=f.device)
        md2 = torch.empty(600, dtype=torch.float64, device=f.device)
        key = torch.zeros(2, dtype=torch.int64, device=f.device)
        ops.kcenter_init(f, 0, 600, cen, md2, key)               # step-wise entry points: per-handle scratch table
        assert int(key[1].item()) == want[0]
        picks, _ = ops.kcenter_greedy(f, [0, 1, 2], 12)
        assert picks.cpu().tolist() == want


def test_device_gaussian_noise_statistics_and_reproducibility():
    """das_add_gaussian_noise (mc_noise.py:26-27 on the device): N(0, sigma) statistics, a pure function of
    (seed, stream id, element), independent streams, ragged / unaligned sizes, in place."""
    ops = _ops()
    n = 1 << 22
    x = torch.zeros(n, device="cuda")
    z = ops.add_gaussian_noise(x, 1.0, seed=1234, stream_id=7)
    m, s = float(z.mean()), float(z.std())
    assert abs(m) < 5 / np.sqrt(n) and abs(s - 1) < 5 / np.sqrt(2 * n)
    zz = z.double()
    assert abs(float((zz ** 3).mean())) < 0.01 and abs(float((zz ** 4).mean()) - 3.0) < 0.02      # skewness 0, kurtosis 3
    assert 4.5 < float(z.abs().max()) < 7.0
    # tail mass like a Gaussian's: P(|z| > 2) = 4.55 %, P(|z| > 3) = 0.27 %
    assert abs(float((z.abs() > 2).float().mean()) - 0.0455) < 0.001 and abs(float((z.abs() > 3).float().mean()) - 0.0027) < 0.0003
    # neighbouring elements / the four outputs of one counter are uncorrelated
    for lag in (1, 2, 3, 4, 1024):
        assert abs(float((z[:-lag] * z[lag:]).mean())) < 6 / np.sqrt(n)
    assert torch.equal(z, ops.add_gaussian_noise(x, 1.0, seed=1234, stream_id=7))                   # reproducible
    z2 = ops.add_gaussian_noise(x, 1.0, seed=1234, stream_id=8)
    z3 = ops.add_gaussian_noise(x, 1.0, seed=1235, stream_id=7)
    for other in (z2, z3):
        assert not torch.equal(z, other) and abs(float((z * other).mean())) < 6 / np.sqrt(n)       # independent
    # x + sigma * z exactly (fma), any size / alignment, in place
    base = torch.randn(1003, device="cuda")
    y = ops.add_gaussian_noise(base, 0.125, seed=5, stream_id=1)
    z1 = ops.add_gaussian_noise(torch.zeros(1003, device="cuda"), 1.0, seed=5, stream_id=1)
    torch.testing.assert_close(y, base + 0.125 * z1, rtol=0, atol=2e-7)
    odd = base[1:1000]                                                   # 4-byte aligned only
    y_odd = ops.add_gaussian_noise(odd.contiguous(), 0.125, seed=5, stream_id=1)
    torch.testing.assert_close(y_odd, odd + 0.125 * z1[:999], rtol=0, atol=2e-7)
    buf = base.clone()
    assert ops.add_gaussian_noise(buf, 0.125, seed=5, stream_id=1, out=buf).data_ptr() == buf.data_ptr()
    assert torch.equal(buf, y)
