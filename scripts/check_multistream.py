"""Graph-replayed training step: side-stream branches vs single stream, and single vs single as the noise floor."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200.train_step import Trainer, default_args

cfg = dict(channel=4, texture_channel=64, N=1, image_size=256, batch_size=2, d_reg_every=4)
g = torch.Generator().manual_seed(11)
batches = [(torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).cuda() for _ in range(4)]


def run(ms):
    torch.manual_seed(21)
    random.seed(21)
    tr = Trainer(default_args(**cfg), device="cuda", seed=9, cuda_graphs=True, multi_stream=ms)
    torch.manual_seed(22)
    random.seed(22)
    out, par = [], None
    for it, X in enumerate(batches, start=1):
        lo = tr.step(X, it)
        out.append({k: float(v) for k, v in lo.items()})
        if it == 1:
            torch.cuda.synchronize()
            par = torch.cat([p.detach().reshape(-1).clone() for k in ("E", "G", "Gstru", "Ex", "Dreal", "Dco", "Ddist")
                             for p in tr.nets[k].parameters()])
    torch.cuda.synchronize()
    return out, par


runs = [run(False), run(False), run(True), run(True)]
names = ["single#1", "single#2", "multi#1", "multi#2"]
for i in range(1, 4):
    a, b = runs[0], runs[i]
    d = (a[1] - b[1]).abs()
    print(f"--- {names[0]} vs {names[i]}: params after it 1: max {float(d.max()):.2e}, frac>1e-4 {float((d > 1e-4).float().mean()):.4f}")
    for it, (x, y) in enumerate(zip(a[0], b[0]), start=1):
        worst = max((abs(x[k] - y[k]) / max(1.0, abs(x[k])), k) for k in x)
        print(f"   it {it}: worst rel loss diff {worst[0]:.2e} ({worst[1]})")
