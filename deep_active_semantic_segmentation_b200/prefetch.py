"""Batch feeder of the selectors: what stands between the caller's dataset and the scoring kernels.

The reference iterates `DataLoader(PathsDataset(...), batch_size, shuffle=False, num_workers=0)` and calls
`.cuda()` on every batch (mc_dropout.py:131-137, ceal.py:21-29, core_set.py:42-56): collate into pageable memory,
then a synchronous staged copy - about 11 ms per batch of eight 512 x 1024 images, invisible next to T network
forwards but ten times the scoring kernels.  `DeviceBatchLoader` keeps the iteration order and the batch boundaries
and shortens the host side:

  * items that already live in PINNED host memory are not touched by the host at all: one cudaMemcpyAsync per item and
    field (plain `copy_(non_blocking=True)` calls on the side stream - ~4 us each, no batched-memcpy API) moves them
    straight from where the dataset keeps them into the device slot.  Measured on the B200 box (8 x 512 x 1024 images +
    labels, 67 MB per batch): 1.4 ms of host memcpy per batch -> 0.07 ms, which is what lets the selector keep up with a
    0.95 ms scoring kernel (`profiles/r2_loader_notes.md`),
  * pageable items are copied into a PINNED staging buffer (one of two alternating slots; the staging buffers are cached
    for the life of the process - page-locking costs ~0.5 ms / MB): `torch.stack(out=)` when the calling thread has an
    intra-op team, otherwise (torchrun sets OMP_NUM_THREADS=1 per rank) copy threads - as many as the rank's share of
    the cores it may run on (os.sched_getaffinity / LOCAL_WORLD_SIZE) - doing one plain memcpy per item and field,
  * one asynchronous copy per field on a side stream into one of two DEVICE slots (cached as well: a fresh
    allocation on a side stream goes through cudaMalloc / cudaFree of the caching allocator and stalled single
    calls by 100-200 ms); the consumer's stream waits on the copy's event; a slot pair is reused only when the device
    has finished the batch that last came out of it, so assembling batch i+1 overlaps the device work of batch i and
    the host never runs more than two batches ahead,
  * batches come out as CUDA tensors (`.cuda()` on them is a no-op), dict fields and bare tensors alike.  A batch is
    a view of a device slot: it stays valid until the loader has been asked for two more batches (the selectors
    consume a batch before they ask for the next one).

The dataset is only ever read on the calling thread, and nothing is assembled ahead of the consumer (threads that
assemble whole batches ahead were measured and dropped: with an intra-op team inside each worker the host is
oversubscribed and the thread that enqueues the GPU work stalls - `profiles/r1_loader_notes.md`).
Datasets whose items are not tensors / arrays of one shape per field fall back to the plain DataLoader;
`DAS_LOADER=torch` forces it.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from torch.utils.data import DataLoader

_STAGING = {}   # (device, field, shape, dtype, batch) -> ([pinned slot 0, pinned slot 1], [device slot 0, device slot 1])
_SIDE = {}      # device -> copy stream
_COPIERS = None  # process-wide copy threads (memcpy releases the GIL)


def copy_threads() -> int:
    """Copy threads of this process: DAS_LOADER_COPY_THREADS, else this rank's share of the cores it may run on
    (torchrun pins nothing but sets OMP_NUM_THREADS=1, so eight ranks would otherwise run eight single-threaded feeders
    on a 16-core host or oversubscribe it), at least 1 and at most 8."""
    env = os.environ.get("DAS_LOADER_COPY_THREADS")
    if env:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover - non-Linux
        cores = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(8, cores // local_world))


def _copiers():
    global _COPIERS
    if _COPIERS is None:
        _COPIERS = ThreadPoolExecutor(max_workers=copy_threads(), thread_name_prefix="das-copy")
    return _COPIERS


def _as_tensor(v):
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, np.ndarray):
        return torch.from_numpy(v)
    raise TypeError(type(v))


def _staging(device, field, t, bs):
    key = (str(device), field, tuple(t.shape), t.dtype, bs)
    if key not in _STAGING:
        if sum(h[0].numel() * h[0].element_size() * 2 for h, _ in _STAGING.values()) > (2 << 30):
            torch.cuda.synchronize()
            _STAGING.clear()          # shapes changed a lot: do not hoard page-locked / device memory
        shape = (bs,) + tuple(t.shape)
        _STAGING[key] = ([torch.empty(shape, dtype=t.dtype, pin_memory=True) for _ in range(2)],
                         [torch.empty(shape, dtype=t.dtype, device=device) for _ in range(2)])
    return _STAGING[key]


def _side_stream(device):
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


class DeviceBatchLoader:
    def __init__(self, dataset, batch_size: int, device=None):
        self.dataset, self.batch_size = dataset, int(batch_size)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        #: host seconds spent assembling / enqueueing batches (dataset reads included; the back-pressure wait for the
        #: device is counted separately in wait_seconds) and the number of batches, of the last iteration - bench.py
        #: reports them as `host_ms_per_batch`
        self.host_seconds, self.wait_seconds, self.batches = 0.0, 0.0, 0
        #: how the fields of the last iteration travelled: "direct" (pinned items, no host copy) or "staged"
        self.path = None

    def __len__(self):
        return -(-len(self.dataset) // self.batch_size)

    def _fallback(self):
        return iter(DataLoader(self.dataset, batch_size=self.batch_size, shuffle=False, num_workers=0))

    def __iter__(self):
        n = len(self.dataset)
        if n == 0:
            return
        if os.environ.get("DAS_LOADER", "") == "torch" or self.device.type != "cuda":
            yield from self._fallback()
            return
        try:
            first = self.dataset[0]
            fields = {k: _as_tensor(v) for k, v in first.items()} if isinstance(first, dict) else {None: _as_tensor(first)}
        except TypeError:
            yield from self._fallback()
            return
        bs = self.batch_size
        slots = {k: _staging(self.device, k, t, bs) for k, t in fields.items()}     # field -> (pinned pair, device pair)
        done = [None, None]     # per slot: the consumer's stream has finished the batch that last came out of it
        side = _side_stream(self.device)
        torch.cuda.current_stream(self.device).synchronize()   # an abandoned earlier loader may still own the slots
        self.host_seconds, self.wait_seconds, self.batches = 0.0, 0.0, 0
        # pinned items go straight to the device slot (one async copy per item and field), pageable ones through the
        # pinned staging slot; decided per field from the first item (cudaPointerGetAttributes, ~1 us)
        direct = {k: bool(t.is_pinned()) for k, t in fields.items()}
        self.path = "direct" if all(direct.values()) else ("staged" if not any(direct.values()) else "mixed")
        for bi, lo in enumerate(range(0, n, bs)):
            t_host = time.perf_counter()
            slot = bi & 1
            cur = torch.cuda.current_stream(self.device)
            if bi > 0:          # the consumer asked for the next batch: everything it enqueued for batch bi-1 is on `cur`
                done[slot ^ 1] = torch.cuda.Event()
                done[slot ^ 1].record(cur)
            items = [first if j == 0 else self.dataset[j] for j in range(lo, min(lo + bs, n))]
            m = len(items)
            if done[slot] is not None:
                # double buffering with back-pressure: batch bi-2 has left the host AND the device is done with it, so
                # the host runs at most two batches ahead and the device tensors of a batch are reused, not re-allocated
                t_wait = time.perf_counter()
                done[slot].synchronize()
                t_wait = time.perf_counter() - t_wait
                self.wait_seconds += t_wait
                t_host += t_wait          # waiting for the device is not host work
            staged = [k for k in slots if not direct[k]]
            if staged and torch.get_num_threads() >= 8:
                # an intra-op team is available on this thread: one stack per field (2 ms per 67 MB batch)
                for k in staged:
                    torch.stack([_as_tensor(it[k] if k is not None else it) for it in items], out=slots[k][0][slot][:m])
            elif staged:
                # torchrun sets OMP_NUM_THREADS=1 per rank: copy threads, one memcpy per item and field (4.5 ms per
                # batch with 4 threads instead of 8 ms for a single-threaded stack)
                jobs = [(slots[k][0][slot][j], _as_tensor(it[k] if k is not None else it))
                        for k in staged for j, it in enumerate(items)]
                for _ in _copiers().map(lambda d_s: d_s[0].copy_(d_s[1]), jobs):
                    pass
            with torch.cuda.stream(side):
                dev = {}
                for k, (pinned, device_) in slots.items():
                    if direct[k]:
                        for j, it in enumerate(items):
                            src = _as_tensor(it[k] if k is not None else it)
                            if not src.is_pinned():       # a later item is pageable after all: stage this one
                                src = pinned[slot][j].copy_(src)
                            device_[slot][j].copy_(src, non_blocking=True)
                        dev[k] = device_[slot][:m]
                    else:
                        dev[k] = device_[slot][:m].copy_(pinned[slot][:m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            cur.wait_event(ev)
            self.host_seconds += time.perf_counter() - t_host
            self.batches += 1
            yield dev[None] if None in dev else dev
            del dev
        torch.cuda.current_stream(self.device).synchronize()   # the cached staging buffers may be reused by the next loader
