#!/usr/bin/env python
"""Host-side cost of ActiveSelectionCoreSet._select_batch at BASELINE config 5 (cProfile of the calling thread)."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_active_semantic_segmentation_b200 import synth
from deep_active_semantic_segmentation_b200.active_selection import ActiveSelectionCoreSet

feats = torch.from_numpy(synth.coreset_features(11, 10000, 2048)).cuda()
cs = ActiveSelectionCoreSet(None, 513, 8)
cs._select_batch(feats, list(range(50)), 8)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    cs._select_batch(feats, list(range(50)), 500)
    torch.cuda.synchronize()
    print("select_batch wall ms", round((time.perf_counter() - t0) * 1e3, 3))
pr = cProfile.Profile(); pr.enable()
cs._select_batch(feats, list(range(50)), 500)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue())
