#!/usr/bin/env python
"""Write tests/golden/ckpt_format.json: the STRUCTURE of a checkpoint saved by the unmodified reference.

Run in the build container only (needs /root/reference):

    TORCH_EXTENSIONS_DIR=/tmp/torch_ext TORCH_CUDA_ARCH_LIST=10.0a python tests/golden/make_ckpt_golden.py

One iteration of the reference's own train.py loop (train.py:21-221) runs on CPU through the App. D shims with
``save_every=1``, so the file is written by the reference's save block (train.py:308-322) itself.  The file
(~0.5 GB: Dreal alone has 26 M parameters whatever ``channel`` is, plus Adam moments) is not committed; what is
committed is its structure -- top-level keys, per-entry state_dict keys / shapes / dtypes, the optimiser
state_dict layout -- which ``ideas_b200.checkpoint.save_checkpoint`` must reproduce and ``load_checkpoint`` must
accept (tests/test_checkpoint.py)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def write_reference_checkpoint(out_dir, ref="/root/reference"):
    """Runs the reference loop for one iteration and returns the path of the checkpoint it saved."""
    import argparse
    import random
    import types

    import torch
    sys.path.insert(0, ref)
    ds = types.ModuleType("dataset")
    ds.set_dataset = lambda **kw: None
    sys.modules["dataset"] = ds
    import models as R
    import utils as RU
    import train as RT
    torch.Tensor.cuda = lambda self, *a, **k: self
    ta = argparse.Namespace(channel=4, structure_channel=8, texture_channel=64, N=1, image_size=256,
                            channel_multiplier=1, blur_kernel=(1, 3, 3, 1), num_iters=1, start_iter=0, lambda_Ex=10.0,
                            lr=0.002, batch_size=1, real_r1=10.0, texture_r1=1.0, dist_r1=1.0, ref_crop=4, n_crop=8,
                            d_reg_every=16, log_every=10 ** 9, show_every=10 ** 9, save_every=1, exp_name="ckpt")
    torch.manual_seed(21)
    random.seed(21)
    order = [("E", "DisentanglementEncoder"), ("G", "Generator"), ("Gstru", "StructureGenerator"),
             ("Ex", "TensorExtractor"), ("Dreal", "ImageLevelDiscriminator"),
             ("Dco", "CooccurenceDiscriminator"), ("Ddist", "DistributionDiscriminator"),
             ("E_ema", "DisentanglementEncoder"), ("G_ema", "Generator"),
             ("Gstru_ema", "StructureGenerator"), ("Ex_ema", "TensorExtractor")]
    trainer = {k: R.init_model(n, ta) for k, n in order}
    for k in ("E", "G", "Gstru", "Ex"):
        trainer[k + "_ema"].eval()
        RU.accumulate(trainer[k + "_ema"], trainer[k], 0)
    P = lambda *ks: [p for k in ks for p in trainer[k].parameters()]  # noqa: E731
    trainer["g_optim"] = torch.optim.Adam(P("E", "G", "Gstru"), lr=ta.lr, betas=(0.0, 0.99))
    trainer["ex_optim"] = torch.optim.Adam(P("Ex"), lr=ta.lr, betas=(0.0, 0.99))
    r = ta.d_reg_every / (ta.d_reg_every + 1)
    trainer["d_optim"] = torch.optim.Adam(P("Dreal", "Dco", "Ddist"), lr=ta.lr * r, betas=(0.0 ** r, 0.99 ** r))
    RT.base_dir = out_dir                                   # globals train() reads when it logs / saves
    RT.ckpt_dir = os.path.join(out_dir, "checkpoints")
    RT.sample_dir = os.path.join(out_dir, "samples")
    os.makedirs(RT.ckpt_dir, exist_ok=True)
    RT.train(exp_name="ckpt", args=ta, loader=[torch.rand(1, 3, 256, 256) * 2 - 1], trainer=trainer, device="cpu")
    return os.path.join(RT.ckpt_dir, "1.pt"), trainer


def describe(ckpt):
    """JSON-able structure of a checkpoint dict."""
    import torch

    def sd_desc(sd):
        return [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]

    out = {"top": sorted(ckpt.keys()), "iter_idx_type": type(ckpt["iter_idx"]).__name__, "N_type": type(ckpt["N"]).__name__,
           "args_type": type(ckpt["args"]).__name__, "args_fields": sorted(vars(ckpt["args"]).keys()), "trainer": {}}
    for k, v in ckpt["trainer"].items():
        if "param_groups" in v:                               # an optimiser state_dict
            st = v["state"]
            first = st[min(st)] if st else {}
            out["trainer"][k] = {"kind": "optim", "n_state": len(st), "state_fields": sorted(first.keys()),
                                 "step_is_tensor": bool(st) and torch.is_tensor(first["step"]),
                                 "group_fields": sorted(v["param_groups"][0].keys()),
                                 "n_params": [len(g["params"]) for g in v["param_groups"]]}
        else:
            out["trainer"][k] = {"kind": "module", "state_dict": sd_desc(v)}
    return out


if __name__ == "__main__":
    import tempfile

    import torch
    with tempfile.TemporaryDirectory() as d:
        path, _ = write_reference_checkpoint(d)
        ck = torch.load(path, map_location="cpu", weights_only=False)
        json.dump(describe(ck), open(os.path.join(HERE, "ckpt_format.json"), "w"), indent=0)
        print("wrote ckpt_format.json from", path, os.path.getsize(path) // 2 ** 20, "MiB")
