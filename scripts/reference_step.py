#!/usr/bin/env python
"""Time the UNMODIFIED reference's training loop (train.py:21-221) on the host cores or on a GPU.

    python scripts/reference_step.py --ref /root/reference --device cpu  --batch 4  --steps 2 --warmup 0
    python scripts/reference_step.py --ref baseline/_ref/IDEAS --device cuda --batch 32 --steps 48 --warmup 16

BENCH INFRASTRUCTURE (the comparator of BASELINE.md §4 / SURVEY.md §8(d) cfg 4) -- never imported by the product.
Nothing in the reference tree is edited; the launcher shims of SURVEY.md App. D live here: a stub `dataset` module
(the reference's needs lmdb / imutils), float Adam betas (train.py:417-426 passes the integer 0), and on the CPU
`Tensor.cuda` -> identity (train.py:61,148 hard-code .cuda()).  The trainer dict is assembled exactly as the
script's __main__ block does (train.py:390-432, cudnn.benchmark = True as train.py:327) and handed to the
reference's own ``train()``; the synthetic loader timestamps every batch fetch (after a device synchronise), so the
difference of two fetches is one iteration.  Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import time
import types


class _Budget(Exception):
    pass


def time_reference(ref, device="cpu", batch=4, image_size=256, steps=2, warmup=0, budget_s=150.0, cudnn_tf32=True,
                   subprocess_ok=True):
    """(seconds per iteration, timed iterations).  Runs in a child process so the reference's module names
    (`models`, `utils`, `stylegan2`) and the Tensor.cuda patch never leak into the caller."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--ref", ref, "--device", device, "--batch", str(batch),
           "--image-size", str(image_size), "--steps", str(steps), "--warmup", str(warmup), "--budget", str(budget_s),
           "--cudnn-tf32", str(int(cudnn_tf32))]
    env = dict(os.environ)
    env.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(os.path.dirname(os.path.abspath(ref)), "_ext"))
    env.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("reference_step failed: " + (r.stderr or r.stdout)[-2000:])
    d = json.loads(lines[-1])
    return d["s_per_step"], d["timed_steps"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--budget", type=float, default=150.0)
    ap.add_argument("--cudnn-tf32", type=int, default=1)
    a = ap.parse_args()

    import torch
    from torch import optim
    ref = os.path.abspath(a.ref)
    sys.path.insert(0, ref)
    ds = types.ModuleType("dataset")
    ds.set_dataset = lambda **kw: None
    sys.modules["dataset"] = ds
    cuda = a.device.startswith("cuda")
    if cuda:
        torch.backends.cudnn.benchmark = True                         # train.py:327
        torch.backends.cudnn.allow_tf32 = bool(a.cudnn_tf32)           # torch default True; matmul tf32 stays off
    else:
        torch.set_num_threads(os.cpu_count() or 1)
        torch.Tensor.cuda = lambda self, *x, **k: self
    import models as R
    import train as RT
    import utils as RU
    args = argparse.Namespace(N=1, lambda_Ex=10.0, lr=0.002, batch_size=a.batch, image_size=a.image_size, real_r1=10.0,
                              texture_r1=1.0, dist_r1=1.0, ref_crop=4, n_crop=8, d_reg_every=16, channel=32,
                              channel_multiplier=1, structure_channel=8, texture_channel=2048, log_every=10 ** 9,
                              show_every=10 ** 9, save_every=10 ** 9, start_iter=0, blur_kernel=(1, 3, 3, 1),
                              num_iters=a.warmup + a.steps, exp_name="bench", ckpt=None)
    torch.manual_seed(0)
    random.seed(0)
    dev = a.device
    order = [("E", "DisentanglementEncoder"), ("G", "Generator"), ("Gstru", "StructureGenerator"),
             ("Ex", "TensorExtractor"), ("Dreal", "ImageLevelDiscriminator"), ("Dco", "CooccurenceDiscriminator"),
             ("Ddist", "DistributionDiscriminator"), ("E_ema", "DisentanglementEncoder"), ("G_ema", "Generator"),
             ("Gstru_ema", "StructureGenerator"), ("Ex_ema", "TensorExtractor")]
    trainer = {k: R.init_model(n, args).to(dev) for k, n in order}
    for k in ("E", "G", "Gstru", "Ex"):
        trainer[k + "_ema"].eval()
        RU.accumulate(trainer[k + "_ema"], trainer[k], 0)
    P = lambda *ks: [p for k in ks for p in trainer[k].parameters()]  # noqa: E731
    trainer["g_optim"] = optim.Adam(P("E", "G", "Gstru"), lr=args.lr, betas=(0.0, 0.99))
    trainer["ex_optim"] = optim.Adam(P("Ex"), lr=args.lr, betas=(0.0, 0.99))
    r = args.d_reg_every / (args.d_reg_every + 1)
    trainer["d_optim"] = optim.Adam(P("Dreal", "Dco", "Ddist"), lr=args.lr * r, betas=(0.0 ** r, 0.99 ** r))

    pool = [torch.rand(a.batch, 3, a.image_size, a.image_size) * 2 - 1 for _ in range(2)]
    if cuda:
        pool = [p.pin_memory() for p in pool]
    stamps = []
    t_begin = time.perf_counter()

    def sync_now():
        if cuda:
            torch.cuda.synchronize()
        return time.perf_counter()

    def loader():
        i = 0
        while True:
            now = sync_now()
            stamps.append(now)
            done = len(stamps) - 1 - a.warmup                          # timed iterations finished so far
            if done >= 1 and now - t_begin + (stamps[-1] - stamps[-2]) > a.budget:
                raise _Budget()
            yield pool[i % 2]
            i += 1

    try:
        RT.train(exp_name="bench", args=args, loader=loader(), trainer=trainer, device=dev)
    except _Budget:
        pass
    else:
        stamps.append(sync_now())
    durs = [b - c for c, b in zip(stamps[:-1], stamps[1:])]
    timed = durs[a.warmup:]
    out = {"impl": "reference", "device": a.device, "batch": a.batch, "image_size": a.image_size,
           "warmup": a.warmup, "timed_steps": len(timed), "s_per_step": sum(timed) / len(timed),
           "images_per_s": a.batch * len(timed) / sum(timed), "ms_per_step_all": [round(d * 1e3, 2) for d in durs],
           "r1_iterations_timed": sum(1 for i in range(a.warmup, len(durs)) if (i + 1) % 16 == 0),
           "threads": torch.get_num_threads(), "torch": torch.__version__}
    if cuda:
        out.update(cudnn_benchmark=True, cudnn_allow_tf32=bool(a.cudnn_tf32), matmul_allow_tf32=torch.backends.cuda.matmul.allow_tf32,
                   peak_mem_gib=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), gpu=torch.cuda.get_device_name(0))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
