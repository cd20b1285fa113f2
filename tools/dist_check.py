#!/usr/bin/env python
"""Multi-GPU parity check (one process per GPU, NCCL): the sharded selectors must return, on every rank, exactly
what the reference's un-sharded classes returned on the same pools (tests/golden/*.npz).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import fakes  # noqa: E402
from tests import golden_util as G  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from deep_active_semantic_segmentation_b200 import constants, dist, synth
    from deep_active_semantic_segmentation_b200.active_selection import base, get_active_selection_class

    base.paths_dataset.PathsDataset = fakes.SyntheticPathsDataset
    paths = lambda n: [str(i) for i in range(n)]
    idx = lambda ps: [int(p) for p in ps]
    RT, AT = 1e-5, 2e-7
    done = []

    for name in ("mc_small", "mc_aligned"):
        g = G.load(name)
        seed, N, T, C, H, W, block, k, bs = (int(v) for v in g["meta"])
        logits, labels = G.pool_from_meta(seed, N, T, C, H, W, block, g["logits_sha"])
        pool = fakes.Pool(logits, labels)
        constants.MC_STEPS = T
        sel = get_active_selection_class("variance", C, pool, H if H == W else -1, bs)
        chosen = sel.get_vote_entropy_for_images(fakes.ReplayModel(pool), paths(N), k)
        assert idx(chosen) == g["ve_selected"].tolist(), (rank, name)
        np.testing.assert_allclose(sel.last_scores, g["ve_scores"], rtol=RT, atol=AT)
        ceal = get_active_selection_class("ceal_margin", C, pool, H if H == W else -1, bs)
        assert idx(ceal.get_least_margin_samples(fakes.ReplayModel(pool), paths(N), k)) == g["ceal_margin_selected"].tolist()
        np.testing.assert_allclose(ceal.last_scores, g["ceal_margin"], rtol=RT, atol=AT)
        ent_sel, ent = ceal.get_maximum_entropy_samples(fakes.ReplayModel(pool), paths(N), k)
        assert idx(ent_sel) == g["ceal_entropy_selected"].tolist()
        done.append(name)

    for name in ("region_small", "region_mid"):
        g = G.load(name)
        seed, N, T, C, S, block, Rg, sel_size, bs = (int(v) for v in g["meta"])
        logits, labels = G.pool_from_meta(seed, N, T, C, S, S, block, g["logits_sha"])
        pool = fakes.Pool(logits, labels)
        constants.MC_STEPS = T
        sel = get_active_selection_class("variance", C, pool, S, bs)
        regions, count = sel.create_region_maps(fakes.ReplayModel(pool), paths(N), G.regions_from_rows(g["existing"], N), Rg, sel_size)
        assert count == int(g["count"]), (rank, name, count)
        assert regions == {str(i): lst for i, lst in enumerate(G.regions_from_rows(g["regions"], N)) if lst}
        done.append(name)

    # fewer images than ranks: ranks with an empty shard must still take part in every exchange
    g = G.load("region_small")
    seed, N, T, C, S, block, Rg, sel_size, bs = (int(v) for v in g["meta"])
    logits, labels = G.pool_from_meta(seed, N, T, C, S, S, block, g["logits_sha"])
    pool = fakes.Pool(logits[:1], labels[:1])
    constants.MC_STEPS = T
    sel = get_active_selection_class("variance", C, pool, S, bs)
    one = sel.create_region_maps(fakes.ReplayModel(pool), paths(1), [[]], Rg, sel_size)
    chosen = sel.get_vote_entropy_for_images(fakes.ReplayModel(pool), paths(1), 3)
    assert chosen == ("0",) and one[1] >= 1 and set(one[0]) == {"0"}, (rank, one, chosen)
    gathered = dist.gather_objects(one)
    assert all(x == gathered[0] for x in gathered)
    done.append("one-image pool")

    for name, force, shard in (("coreset_small", True, True), ("coreset_mid", True, True), ("coreset_mid", False, True),
                               ("coreset_mid", "auto", "auto")):
        g = G.load(name)
        seed, N, D, L, K = (int(v) for v in g["meta"])
        feats = synth.coreset_features(seed, N, D)
        sel = get_active_selection_class("coreset", 2, None, None, 4)
        sel.tensor_core_filter = force           # per-rank tcgen05 filter on / off
        sel.shard_rows = shard                   # rows of min_d sharded (one exchange per step) / replicated loop
        picks = sel._select_batch(feats.astype(np.float64), list(range(L)), K)
        assert picks == g["picks"].tolist(), (rank, name)
        np.testing.assert_allclose(sel.last_min_distances.cpu().numpy(), g["min_dist"], rtol=1e-9, atol=1e-4)
        done.append(f"{name}/filter={force}/shard={shard}")

    td.barrier()
    if rank == 0:
        print(f"[dist_check] world={world} ok: {', '.join(done)}", flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
