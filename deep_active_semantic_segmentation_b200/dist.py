"""Pool sharding over ranks (one process per GPU) and the tiny exchanges the path needs.

The unlabeled pool shards by image: rank r scores the contiguous block
[r*ceil(N/W), min(N, (r+1)*ceil(N/W))) with no communication.  The only exchanges are
  * an all-gather of each rank's top-k (score, global index) candidates,
  * an all-reduce of the pool min / max of the region score maps (mc_dropout.py:152-153),
  * an all-gather of the per-image NMS pick sequences,
  * for core-set, one tiny all-gather per greedy step (the arg-max key).
Everything below works with any torch.distributed backend: NCCL with CUDA tensors on the GPU box,
gloo with CPU tensors in the CPU tests.  Without an initialised process group it degrades to W = 1.
"""
from __future__ import annotations

import heapq
import math

import torch
import torch.distributed as td


def world():
    if td.is_available() and td.is_initialized():
        return td.get_world_size(), td.get_rank()
    return 1, 0


def _comm_device():
    """NCCL moves CUDA tensors, gloo moves CPU tensors."""
    if td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPU cores of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node
    and that node's cpulist), so that the pinned staging buffers it allocates afterwards are first-touched on that node
    and its copy threads run next to them.  torchrun starts one process per GPU without any binding: on a two-socket
    box half of the ranks then feed their GPU across the socket interconnect.  Returns what was done (for logs);
    does nothing when the topology is not visible (containers without sysfs, single-node machines)."""
    import os

    info = {"device": int(device_index), "bound": False}
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        sysdev = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        with open(os.path.join(sysdev, "numa_node")) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:  # topology not visible: run unbound, as before
        info["error"] = repr(exc)[:120]
    return info


def shard_bounds(n: int, world_size: int, rank: int):
    per = -(-n // world_size) if n > 0 else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_candidates(scores, ids, k: int):
    """All-gather up to k (score, global id) candidates per rank -> python lists over all ranks.

    scores/ids: 1-D tensors (any device) already restricted to the rank's best k, in rank-local order.
    """
    W, _ = world()
    s = scores.detach().to(torch.float32).reshape(-1)
    i = ids.detach().to(torch.int64).reshape(-1)
    if W == 1:
        return s.cpu().tolist(), i.cpu().tolist()
    dev = _comm_device()
    buf = torch.zeros(k, 2, dtype=torch.float64, device=dev)     # (score, id) - ids < 2^53 are exact
    cnt = torch.tensor([s.numel()], dtype=torch.int64, device=dev)
    if s.numel():
        buf[: s.numel(), 0] = s.to(dev, torch.float64)
        buf[: s.numel(), 1] = i.to(dev, torch.float64)
    bufs = [torch.empty_like(buf) for _ in range(W)]
    cnts = [torch.empty_like(cnt) for _ in range(W)]
    td.all_gather(bufs, buf)
    td.all_gather(cnts, cnt)
    out_s, out_i = [], []
    for b, c in zip(bufs, cnts):
        n = int(c.item())
        b = b[:n].cpu()
        out_s += b[:, 0].tolist()      # float32 values widened exactly
        out_i += [int(v) for v in b[:, 1].tolist()]
    return out_s, out_i


def exchange_records(rec: torch.Tensor) -> torch.Tensor:
    """ONE all_gather_into_tensor of every rank's [k, 2] int64 candidate-record block (ops.topk_records) -> the
    [W * k, 2] table in rank order, on the device the block came from.  Works with NCCL (CUDA) and gloo (CPU)."""
    W, _ = world()
    if W == 1:
        return rec
    dev = _comm_device()
    send = rec.detach().to(dev).contiguous()
    recv = torch.empty((W * send.shape[0], 2), dtype=torch.int64, device=dev)
    td.all_gather_into_tensor(recv, send)
    return recv.to(rec.device)


def select_ranked(scores: torch.Tensor, k: int, descending: bool, ids: torch.Tensor | None = None, id_offset: int = 0):
    """First k of the pool-global stable ranking from each rank's local scores: K3 on the shard (records written
    straight into the all-gather send buffer) -> one all-gather -> K3 merge on every rank -> ONE device-to-host copy of
    the k winners.  Returns (float32 scores, int64 global ids) numpy arrays, identical on all ranks; fewer than k when
    the pool is smaller.  ids: per-candidate int64 payload (region flat indices), default = position + id_offset (the
    global image index of a contiguous shard).  Ties rank by (rank, local order) = global index ascending."""
    from . import ops

    W, _ = world()
    rec = ops.topk_records(scores, int(k), descending, ids=ids, id_offset=id_offset)
    if W > 1:
        rec = ops.topk_merge(exchange_records(rec), int(k), descending)
    return ops.records_to_host(rec)


def merge_ranked(scores, ids, k: int, descending: bool):
    """Stable global ranking of gathered candidates: score, then global index ascending - the order
    Python's stable sorted() gives on the un-sharded pool (mc_dropout.py:195, ceal.py:69)."""
    order = sorted(range(len(scores)), key=(lambda j: (-scores[j], ids[j])) if descending else (lambda j: (scores[j], ids[j])))
    order = order[:k]
    return [scores[j] for j in order], [ids[j] for j in order]


def gather_ranked_np(scores, ids, k: int, descending: bool = True):
    """numpy twin of gather_candidates + merge_ranked for large k (region candidates): every rank's head of at
    most k (score, global id) pairs, already sorted, -> the first k of their stable merge as (float32 [m],
    int64 [m]) arrays.  Order: score, then global id ascending (the reference's first-arg-max tie rule)."""
    import numpy as np

    W, _ = world()
    s = scores.detach().to(torch.float32).reshape(-1)
    i = ids.detach().to(torch.int64).reshape(-1)
    if W == 1:
        return s.cpu().numpy()[:k], i.cpu().numpy()[:k]      # one rank: K3 already produced this order
    dev = _comm_device()
    buf_s = torch.full((k,), float("-inf") if descending else float("inf"), dtype=torch.float32, device=dev)
    buf_i = torch.full((k,), -1, dtype=torch.int64, device=dev)
    buf_s[: s.numel()] = s.to(dev)
    buf_i[: i.numel()] = i.to(dev)
    all_s = [torch.empty_like(buf_s) for _ in range(W)]
    all_i = [torch.empty_like(buf_i) for _ in range(W)]
    td.all_gather(all_s, buf_s)
    td.all_gather(all_i, buf_i)
    gs = torch.cat(all_s).cpu().numpy()
    gi = torch.cat(all_i).cpu().numpy()
    keep = gi >= 0
    gs, gi = gs[keep], gi[keep]
    order = np.lexsort((gi, -gs if descending else gs))[:k]
    return gs[order], gi[order]


def allreduce_minmax(minmax: torch.Tensor) -> torch.Tensor:
    """Pool-global min / max from the per-rank (min, max) pair."""
    W, _ = world()
    if W == 1:
        return minmax
    dev = _comm_device()
    t = torch.stack([-minmax[0], minmax[1]]).to(dev)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return torch.stack([-t[0], t[1]]).to(minmax.device)


def all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Row blocks of a [n_total, D] matrix that is sharded by `shard_bounds` -> the whole matrix on every rank
    (core-set: each rank forwards its slice of the images, core_set.py:56-63, then all ranks run the same greedy loop).
    One all_gather_into_tensor of equally sized (padded) blocks; a rank with an empty shard contributes padding only."""
    W, rank = world()
    if W == 1:
        return local
    per = -(-n_total // W) if n_total > 0 else 0
    dev = _comm_device()
    d = torch.tensor([local.shape[1] if local.dim() == 2 and local.shape[0] > 0 else 0], dtype=torch.int64, device=dev)
    td.all_reduce(d, op=td.ReduceOp.MAX)
    D = int(d.item())
    send = torch.zeros((per, D), dtype=local.dtype, device=dev)
    if local.shape[0] > 0:
        send[: local.shape[0]] = local.to(dev)
    recv = torch.empty((W * per, D), dtype=local.dtype, device=dev)
    td.all_gather_into_tensor(recv, send)
    parts = []
    for r in range(W):
        lo, hi = shard_bounds(n_total, W, r)
        parts.append(recv[r * per: r * per + (hi - lo)])
    return torch.cat(parts).to(local.device).contiguous()


def gather_objects(obj):
    """All-gather a small picklable object (per-image NMS sequences) -> list over ranks."""
    W, _ = world()
    if W == 1:
        return [obj]
    out = [None] * W
    td.all_gather_object(out, obj)
    return out


def sequences_from_device(cand_score, cand_rc, cand_count):
    """(scores [N,kmax], rc [N,kmax,2], count [N]) device tensors of das_nms_sequences -> list over images of
    [(score, r, c), ...] in pick order.  One D2H copy per tensor, no per-candidate synchronisation."""
    cs = cand_score.detach().cpu().numpy()
    rc = cand_rc.detach().cpu().numpy()
    cnt = cand_count.detach().cpu().numpy()
    return [[(cs[i, j], int(rc[i, j, 0]), int(rc[i, j, 1])) for j in range(int(cnt[i]))] for i in range(cs.shape[0])]


def merge_nms_sequences(seqs, region_size: int, max_selection_count: float, H2: int, W2: int, stop: float = 0.01):
    """Global greedy NMS == k-way merge of the image-local pick sequences.

    seqs: list over images (global order) of [(score, r, c), ...] in pick order.
    Order: score descending, ties by flat pool index (image, r, c) ascending = the first-flat-argmax
    rule of mc_dropout.py:91.  Pick j+1 is taken iff j < ceil(K) and its score - the pool maximum
    after pick j - is >= 0.01 (mc_dropout.py:87,105); the first pick is unconditional.
    """
    import numpy as np

    stop32 = np.float32(stop)
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
        if count > 0 and np.float32(-negs) < stop32:
            break
        _, r, c = seqs[i][j]
        selected[i].append((int(r), int(c), region_size, region_size))
        count += 1
        if j + 1 < len(seqs[i]):
            s2, r2, c2 = seqs[i][j + 1]
            heapq.heappush(heap, (-float(s2), (i * H2 + r2) * W2 + c2, i, j + 1))
    return selected, count
