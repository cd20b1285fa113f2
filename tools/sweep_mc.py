"""GPU exploration: time the MC scoring step (T passes of K1 + K2) over batch size B and pass-group G.
Prints one line per configuration: images/s, algorithmic GB/s of K1 and fraction of the measured peak.

    python tools/sweep_mc.py [cs|pascal] [--iters 20]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deep_active_semantic_segmentation_b200 import ops, synth  # noqa: E402

SHAPES = {"cs": (512, 1024, 19, 20), "pascal": (513, 513, 21, 20)}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cs"
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 20
    H, W, C, T = SHAPES[which]
    peak = 6548.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    dev = torch.device("cuda", 0)
    rows = []
    for B in (1, 2, 4, 8):
        passes, labels = synth.device_pass_logits(1, 0, B, T, C, H, W, dev)
        for votes, probs in ((True, True), (True, False)):
            for G in (1, 2, 4, 5, 10, 20):
                st = ops.MCState(B, C, H, W, T, votes=votes, probs=probs, device=dev, single_shot=(G >= T))
                groups = [passes[t0:t0 + G] for t0 in range(0, T, G)]

                def step():
                    st.reset()
                    for g in groups[:-1]:
                        st.accumulate(g)
                    st.score(groups[-1], labels, maps=())

                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                # K1 alone
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(iters):
                    st.reset()
                    for g in groups[:-1]:
                        st.accumulate(g)
                a1.record()
                torch.cuda.synchronize()
                ms_acc = a0.elapsed_time(a1) / iters
                gb = T * B * C * H * W * 4 / 1e9
                row = dict(shape=which, B=B, G=G, votes=votes, probs=probs, ms_step=round(ms, 4), ms_k1=round(ms_acc, 4),
                           img_s=round(B / ms * 1e3, 1), k1_gbs=round(gb / ms_acc * 1e3, 1),
                           k1_frac=round(gb / ms_acc * 1e3 / peak, 3), step_frac=round(gb / ms * 1e3 / peak, 3))
                rows.append(row)
                print(json.dumps(row), flush=True)
        del passes
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
