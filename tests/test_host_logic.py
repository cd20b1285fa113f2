"""CPU: the C-ABI library loads and exports every declared symbol (no compute calls without a GPU),
the header and the ctypes table agree, and the host-side sharding / merge logic is correct -
including a world_size-2 gloo run of the exchange steps."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import restate as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    import ctypes
    from deep_active_semantic_segmentation_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "das_b200.h")).read()
    declared = set(re.findall(r"\b(das_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)
    assert lib.das_abi_version() == _lib.ABI_VERSION == 4
    assert lib.das_strerror(0) == b"ok" and b"invalid" in lib.das_strerror(-1)
    # argument validation happens before any CUDA call, so these are safe without a device
    nbytes = ctypes.c_size_t()
    d = _lib.McDesc(2, 19, 512, 1024, 20, 3)
    assert lib.das_mc_state_bytes(ctypes.byref(d), ctypes.byref(nbytes)) == 0
    assert nbytes.value >= 2 * (19 + 1) * 512 * 1024 * 4 + 2 * 20 * 512 * 1024
    assert lib.das_mc_state_bytes(ctypes.byref(_lib.McDesc(1, 33, 8, 8, 4, 3)), ctypes.byref(nbytes)) == -2
    assert lib.das_mc_state_bytes(ctypes.byref(_lib.McDesc(1, 5, 8, 8, 300, 3)), ctypes.byref(nbytes)) == -2
    assert lib.das_mc_state_bytes(ctypes.byref(_lib.McDesc(0, 5, 8, 8, 3, 3)), ctypes.byref(nbytes)) == -1
    assert lib.das_mc_state_bytes(ctypes.byref(_lib.McDesc(1, 5, 8, 8, 3, 0)), ctypes.byref(nbytes)) == -1
    assert lib.das_box_sum_workspace_bytes(1, 16, 16, 17, ctypes.byref(nbytes)) == -1     # R > H
    assert lib.das_topk(None, None, None, 10, 3, 1, None, None, None, None) == -1
    # a NULL (or foreign) handle is refused before anything is enqueued
    assert lib.das_minmax_init(None, None, None) == -1 and lib.das_handle_device(None) == -1
    assert lib.das_handle_set_option(None, 0, 1) == -1 and lib.das_handle_destroy(None) == 0


def test_product_path_has_no_cpu_fallback():
    from deep_active_semantic_segmentation_b200 import ops
    from deep_active_semantic_segmentation_b200._lib import DasError
    with pytest.raises(DasError):
        ops.topk(torch.zeros(8), 2, True)
    with pytest.raises(DasError):
        ops.box_sum(torch.zeros(1, 8, 8), 2, torch.zeros(2))
    # nothing under the package imports the oracle
    pkg = os.path.join(ROOT, "deep_active_semantic_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle can use", ""), f


def test_shard_bounds_cover_the_pool_exactly():
    from deep_active_semantic_segmentation_b200 import dist
    for n in (0, 1, 7, 8, 2975, 10582):
        for W in (1, 2, 4, 8):
            spans = [dist.shard_bounds(n, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi - lo <= -(-n // W) for lo, hi in spans)
    assert dist.shard_bounds(10582, 8, 7) == (9261, 10582)


def test_merge_ranked_equals_unsharded_stable_sort():
    from deep_active_semantic_segmentation_b200 import dist
    rng = np.random.default_rng(0)
    for trial in range(20):
        n, W, k = int(rng.integers(1, 200)), int(rng.integers(1, 9)), int(rng.integers(1, 40))
        scores = np.round(rng.random(n), 1 if trial % 2 else 6).astype(np.float32).tolist()
        for desc in (True, False):
            cand_s, cand_i = [], []
            for r in range(W):
                lo, hi = dist.shard_bounds(n, W, r)
                loc = R.rank_topk(scores[lo:hi], k, desc)            # stands in for the per-rank K3 kernel
                cand_s += [scores[lo + j] for j in loc]
                cand_i += [lo + j for j in loc]
            _, ids = dist.merge_ranked(cand_s, cand_i, min(k, n), desc)
            assert ids == R.rank_topk(scores, k, desc)


def test_merge_nms_sequences_equals_reference_loop():
    from deep_active_semantic_segmentation_b200 import dist
    rng = np.random.default_rng(1)
    for trial in range(8):
        N, H2, W2, Rg = int(rng.integers(1, 6)), int(rng.integers(6, 30)), int(rng.integers(6, 30)), int(rng.integers(2, 7))
        m = rng.random((N, H2, W2)).astype(np.float32)
        if trial % 2:
            m = np.round(m, 1)
        K = float(rng.integers(1, 30)) + 0.25
        want = R.square_nms(m.copy(), Rg, K)
        seqs = [R.nms_sequence_single(m[i].copy(), Rg, int(np.ceil(K))) for i in range(N)]
        assert dist.merge_nms_sequences(seqs, Rg, K, H2, W2) == want


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as td
from deep_active_semantic_segmentation_b200 import dist
from oracle import restate as R
td.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
W, rank = dist.world()
assert (W, rank) == (2, int(sys.argv[3]))
rng = np.random.default_rng(4)
n, k = 41, 9
scores = np.round(rng.random(n), 1).astype(np.float32)
lo, hi = dist.shard_bounds(n, W, rank)
for desc in (True, False):
    loc = R.rank_topk(scores[lo:hi].tolist(), k, desc)
    s, i = dist.gather_candidates(torch.tensor(scores[lo:hi][loc]), torch.tensor(loc) + lo, k)
    _, ids = dist.merge_ranked(s, i, k, desc)
    assert ids == R.rank_topk(scores.tolist(), k, desc), (ids, desc)
# pool min/max of the region score maps
mm = dist.allreduce_minmax(torch.tensor([1.0 + rank, 5.0 - rank]))
assert mm.tolist() == [1.0, 5.0]
# NMS sequences: every rank ends with the same global selection as the sequential loop
m = np.round(rng.random((5, 12, 14)), 1).astype(np.float32)
want = R.square_nms(m.copy(), 3, 11.5)
lo, hi = dist.shard_bounds(5, W, rank)
local = [R.nms_sequence_single(m[i].copy(), 3, 12) for i in range(lo, hi)]
seqs = [s for part in dist.gather_objects(local) for s in part]
assert dist.merge_nms_sequences(seqs, 3, 11.5, 12, 14) == want
# the same selection through the flattened candidate table (what the device path does: local top-k over
# (score, flat pool index) pairs -> all-gather of the heads -> lexsort -> stop rule)
H2, W2, kmax, K = 12, 14, 12, 11.5
cand_s = np.full((hi - lo, kmax), -np.inf, np.float32)
cand_f = np.full((hi - lo, kmax), -1, np.int64)
for a, seq in enumerate(local):
    for j, (sc, r, c) in enumerate(seq):
        cand_s[a, j], cand_f[a, j] = sc, ((lo + a) * H2 + r) * W2 + c
order = np.lexsort((np.arange(cand_s.size), -cand_s.reshape(-1)))[:12]          # stable: ties keep table order
gs, gi = dist.gather_ranked_np(torch.tensor(cand_s.reshape(-1)[order]), torch.tensor(cand_f.reshape(-1)[order]), 12, True)
ok = (gi >= 0) & np.isfinite(gs)
ok[1:] &= gs[1:] >= np.float32(0.01)
count = int(len(ok) if ok.all() else np.argmin(ok))
regions = [[] for _ in range(5)]
for fid in gi[:count].tolist():
    img, rem = divmod(fid, H2 * W2)
    regions[img].append((rem // W2, rem % W2, 3, 3))
assert (regions, count) == want, (regions, count, want)
# candidate RECORDS (the device path: das_topk_records -> exchange_records -> das_topk_merge), here packed and merged
# by numpy stand-ins with the same layout: int64 pairs {float32 bits, global id}, padding {-/+inf, -1}
def pack(sc, ids, k, desc):
    rec = np.empty((k, 2), np.int64)
    rec[:, 0] = np.array([-np.inf if desc else np.inf], np.float32).view(np.uint32)[0]
    rec[:, 1] = -1
    rec[:len(sc), 0] = np.asarray(sc, np.float32).view(np.uint32)
    rec[:len(sc), 1] = ids
    return torch.from_numpy(rec)
lo, hi = dist.shard_bounds(n, W, rank)
for desc in (True, False):
    loc = R.rank_topk(scores[lo:hi].tolist(), k, desc)
    table = dist.exchange_records(pack(scores[lo:hi][loc], np.array(loc) + lo, k, desc))
    assert table.shape == (W * k, 2) and table.dtype == torch.int64
    t = table.numpy()
    sc = t[:, 0].astype(np.uint32).view(np.float32)
    real = t[:, 1] >= 0
    order = np.lexsort((np.arange(len(sc)), ~real, -sc if desc else sc))    # padding last, ties keep table order
    got = t[order[:k], 1].tolist()
    assert got == R.rank_topk(scores.tolist(), k, desc), (got, desc)
# feature rows of a sharded forward pass (core-set): every rank ends with the whole matrix, ragged last shard included
full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
lo, hi = dist.shard_bounds(7, W, rank)
assert torch.equal(dist.all_gather_rows(full[lo:hi], 7), full)
one = torch.ones(1, 3)
assert torch.equal(dist.all_gather_rows(one[:1] if rank == 0 else one[:0], 1), one)
td.barrier()
print("RANK_OK", rank)
'''


def test_gloo_world_size_2_exchanges(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {r}" in out, out


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: one JSON line with the
    contract keys, same metric / unit / workload as the B200 arm."""
    import json
    env = dict(os.environ, DAS_REF_BUDGET_S="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("pool images scored/sec") and d["value"] > 0 and d["steps"] == 2
    assert d["config"]["workload"] == "mc_dropout_entropy_bald_cityscapes_pool_512x1024_c19_t20"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_header_is_plain_c(tmp_path):
    """include/das_b200.h is the C ABI: it must compile as C99 with no torch / C++ types in any signature."""
    src = tmp_path / "h.c"
    src.write_text('#include "das_b200.h"\nint main(void) { return DAS_ABI_VERSION + DAS_N_SCORES + DAS_ACC_N; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    header = open(os.path.join(ROOT, "include", "das_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)              # declarations only, comments stripped
    assert "torch" not in code and "std::" not in code and "at::" not in code and "Tensor" not in code


def test_upsample_shape_support_is_decided_on_the_host():
    """das_mc_upsample_supported is host-only code (no CUDA call): the window check replays the kernel's float
    arithmetic.  DeepLab's stride-4 decoder shapes and Fast-SCNN's factor 8 are in range, factor 2 is not."""
    from deep_active_semantic_segmentation_b200 import _lib
    lib = _lib.load()
    ok = lambda h, w, H, W: bool(lib.das_mc_upsample_supported(None, h, w, H, W))
    assert ok(128, 256, 512, 1024) and ok(129, 129, 513, 513) and ok(17, 17, 65, 65) and ok(12, 16, 48, 64)
    assert ok(64, 128, 512, 1024)            # models/fastscnn.py:22 (classifier at 1/8)
    assert not ok(32, 32, 64, 64) and not ok(64, 64, 64, 64) and not ok(100, 100, 300, 300)
    assert not ok(0, 4, 16, 16) and not ok(4, 4, 0, 16)
    assert ok(1, 1, 16, 16) and ok(5, 5, 16, 16)


def test_upsample_kernel_variant_is_chosen_by_class_count_width_and_flags():
    """das_mc_upsample_variant (host only): the measured choice of csrc/mc_api.cu up_warps() - one pixel per lane with 20 + 4
    warps up to 20 classes, pixel pairs at 21, one pixel per lane with 16 + 4 warps from 22 on (the pixel-pair kernel
    spills there); vote-only scoring and narrow outputs stay with the pixel-pair kernels; unsupported factors give 0."""
    from deep_active_semantic_segmentation_b200 import ops
    v = lambda B, C, h, w, H, W, **kw: ops.upsample_variant(B, C, h, w, H, W, default_options=True, **kw)
    assert v(8, 19, 128, 256, 512, 1024) == 220 and v(8, 20, 129, 129, 513, 513) == 220 and v(8, 2, 128, 256, 512, 1024) == 220
    assert v(8, 21, 129, 129, 513, 513) == 15 and v(8, 21, 128, 256, 512, 1024) == 15
    assert v(8, 22, 128, 256, 512, 1024) == 216 and v(4, 32, 129, 129, 513, 513) == 216
    assert v(8, 19, 128, 256, 512, 1024, probs=False) == 15 and v(8, 32, 128, 256, 512, 1024, probs=False) == 15
    assert v(8, 19, 128, 256, 512, 1024, votes=False) == 220
    assert v(2, 19, 12, 16, 48, 64) == 4 and v(2, 21, 17, 17, 65, 65) == 4 and v(2, 24, 5, 5, 16, 16) == 4   # narrow outputs
    assert v(2, 19, 17, 20, 65, 80) == 220 and v(2, 19, 17, 20, 65, 79) == 4     # two 40-column tiles per row or not
    assert v(2, 24, 17, 17, 65, 64) == 216                                       # 32-column tiles from 22 classes on
    assert v(2, 19, 32, 32, 64, 64) == 0 and v(2, 40, 128, 256, 512, 1024) == 0 and v(0, 19, 128, 256, 512, 1024) == 0
    # source offsets of the whole batch must fit 32 bits in the one-pixel kernel: 2^32 / (19 * 128 * 256) = 6898.3 images
    assert v(6898, 19, 128, 256, 512, 1024) == 220 and v(6899, 19, 128, 256, 512, 1024) == 15


def test_align_corners_axis_of_the_oracle():
    from oracle import restate as R
    for n_in, n_out in ((128, 512), (129, 513), (256, 1024), (5, 16), (1, 7), (9, 9)):
        i0, i1, l0, l1 = R._align_corners_axis(n_in, n_out)
        assert i0.min() == 0 and i1.max() <= n_in - 1 and ((i1 - i0 == 0) | (i1 - i0 == 1)).all()
        assert (l0 >= 0).all() and (l1 >= 0).all() and np.allclose(l0 + l1, 1.0)
        assert (np.diff(i0) >= 0).all()
        if n_out > 1:
            assert i0[-1] + (l1[-1] > 0.5) == n_in - 1   # the last output sample sits on the last input sample


def test_batch_feeder_falls_back_to_the_dataloader_without_cuda(monkeypatch):
    """prefetch.DeviceBatchLoader on a box without a GPU (or with DAS_LOADER=torch) is the reference's DataLoader loop"""
    from deep_active_semantic_segmentation_b200.prefetch import DeviceBatchLoader

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 5

        def __getitem__(self, i):
            return {"image": torch.full((3, 4, 4), float(i)), "label": torch.full((4, 4), float(-i))}

    monkeypatch.setenv("DAS_LOADER", "torch")
    batches = list(DeviceBatchLoader(DS(), 2, device="cpu"))
    assert [b["image"].shape[0] for b in batches] == [2, 2, 1] and len(DeviceBatchLoader(DS(), 2, device="cpu")) == 3
    assert float(batches[2]["label"][0, 0, 0]) == -4.0 and not batches[0]["image"].is_cuda
    assert list(DeviceBatchLoader(torch.utils.data.TensorDataset(torch.zeros(0, 1)), 2, device="cpu")) == []


def test_selections_txt_is_what_the_reference_tools_read_and_write(tmp_path):
    """utils/saver.py:68-78 format, diffed by the reference's own utils/compare_selections.py when the reference tree
    is present (this container), and by the in-package twin everywhere."""
    from deep_active_semantic_segmentation_b200 import selections
    from oracle import restate as R
    from tests import golden_util as G

    g = G.load("mc_small")
    k = int(g["meta"][7])
    ref_paths = [str(i).encode() for i in g["ve_selected"].tolist()]                      # what the reference selected
    own_paths = [str(i).encode() for i in R.rank_topk(g["ve_scores"].tolist(), k, True)]  # restated ranking of its scores
    r = G.load("region_small")
    ref_regions = {}
    for i, rr, cc, hh, ww in r["regions"].tolist():
        ref_regions.setdefault(str(i).encode(), []).append((rr, cc, hh, ww))
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    for folder, paths in ((a, ref_paths), (b, own_paths)):
        f0 = selections.write_selections(folder, paths, run=0)
        selections.write_selections(folder, list(ref_regions), ref_regions, run=1)
        assert open(f0).read() == "".join(p.decode() + "\n" for p in paths)
    line = open(os.path.join(a, "run_0001", "selections.txt")).readline().strip().split(",")
    first = next(iter(ref_regions))
    assert line[0] == first.decode() and [int(v) for v in line[1:5]] == list(ref_regions[first][0])
    assert (len(line) - 1) % 4 == 0
    assert selections.compare_selections(a, b) == [("run_0000", k, k), ("run_0001", len(ref_regions), len(ref_regions))]
    # the saver of the reference writes byte-identical files
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):
        out = subprocess.run([sys.executable, os.path.join(ref_root, "utils", "compare_selections.py"), a, b],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        assert f"run_0000 = {k}/{k} (100.0)" in out.stdout and "run_0001" in out.stdout
    # a diverging selection is reported, an unequal count is an error
    selections.write_selections(b, own_paths[:-1] + [b"999"], run=0)
    assert selections.compare_selections(a, b)[0] == ("run_0000", k - 1, k)
    selections.write_selections(b, own_paths[:-1], run=0)
    with pytest.raises(ValueError):
        selections.compare_selections(a, b)


def test_numa_binding_helper_is_harmless_without_topology():
    """dist.bind_to_gpu_numa_node never raises: without a visible PCI / NUMA topology (this container: no GPU at all) it
    reports why and leaves the affinity of the process alone."""
    from deep_active_semantic_segmentation_b200 import dist
    before = os.sched_getaffinity(0)
    info = dist.bind_to_gpu_numa_node(0)
    assert info["device"] == 0 and isinstance(info["bound"], bool)
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
    from deep_active_semantic_segmentation_b200 import prefetch
    assert 1 <= prefetch.copy_threads() <= 8
