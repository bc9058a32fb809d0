"""Iteration-1 G_rec_loss = l1(G(E(X)), X) depends on nothing but the initial weights and X: print it for eager / graph x
single / multi stream x concurrent generator on / off to see which configuration computes something else."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200.train_step import Trainer, default_args

cfg = dict(channel=int(os.environ.get("CH", "4")), texture_channel=64, N=1, image_size=256, batch_size=2, d_reg_every=4)
g = torch.Generator().manual_seed(11)
batches = [(torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).cuda() for _ in range(3)]
ONLY = os.environ.get("ONLY", "0") == "1"
for graphs in ((True,) if ONLY else (False, True)):
    for ms in ((True,) if ONLY else (False, True)):
        for conc in ((True,) if ONLY else ((False, True) if ms else (False,))):
            torch.manual_seed(21)
            random.seed(21)
            tr = Trainer(default_args(**cfg), device="cuda", seed=9, cuda_graphs=graphs, multi_stream=ms, concurrent_generator=conc,
                         split_dreal=False)
            torch.manual_seed(22)
            random.seed(22)
            out = []
            for it, X in enumerate(batches, start=1):
                lo = tr.step(X, it)
                out.append((float(lo["G_rec_loss"]), float(lo["E_dist_loss"]), float(lo["D_dist_loss"])))
            torch.cuda.synchronize()
            print(f"graphs={graphs} multi_stream={ms} concurrent_g={conc}: " + "  ".join(f"({a:.6f},{b:.6f},{c:.6f})" for a, b, c in out), flush=True)
            del tr
