"""Checkpoints in the reference's on-disk format (reference train.py:308-322, :435-442).

A reference checkpoint is ``torch.save({'iter_idx': int, 'N': int, 'trainer': {key: state_dict}, 'args':
argparse.Namespace})`` with one ``state_dict`` per entry of the trainer dict: the seven networks, the four EMA
copies and the three Adam optimisers.  Module names, parameter / buffer keys and shapes of the B200 networks are
the reference's (tests/golden/contract.json), so the files are interchangeable in both directions: a checkpoint
written by the reference resumes here, and one written here resumes in the reference.

Tensors are saved on the CPU (the reference saves CUDA tensors and maps them back at load with
``map_location``, train.py:437; either loads in either implementation).
"""
from __future__ import annotations

import argparse
from typing import Optional

import torch

TRAINER_KEYS = ("E", "G", "Gstru", "Ex", "Dreal", "Dco", "Ddist", "E_ema", "G_ema", "Gstru_ema", "Ex_ema",
                "g_optim", "ex_optim", "d_optim")


def _cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


def checkpoint_dict(trainer, iter_idx: int, args: Optional[argparse.Namespace] = None) -> dict:
    """The dictionary train.py:309-320 saves, built from a ``Trainer`` (or any mapping with the reference's keys)."""
    args = args if args is not None else trainer.args
    return {"iter_idx": int(iter_idx), "N": int(args.N),
            "trainer": {k: _cpu(trainer[k].state_dict()) for k in TRAINER_KEYS}, "args": args}


def save_checkpoint(trainer, path: str, iter_idx: int, args: Optional[argparse.Namespace] = None) -> None:
    torch.save(checkpoint_dict(trainer, iter_idx, args), path)


def load_checkpoint(trainer, path: str, strict: bool = True) -> int:
    """Resume as train.py:435-442 does: every entry of the trainer dict takes its state_dict; returns the stored
    ``iter_idx`` (the caller's ``start_iter``).  The file holds a pickled Namespace, hence ``weights_only=False``
    -- load only checkpoints you trust."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in TRAINER_KEYS if k not in ckpt["trainer"]]
    if missing and strict:
        raise KeyError(f"checkpoint {path!r} lacks trainer entries {missing}")
    for k in TRAINER_KEYS:
        if k in ckpt["trainer"]:
            trainer[k].load_state_dict(ckpt["trainer"][k])
    return int(ckpt["iter_idx"])
