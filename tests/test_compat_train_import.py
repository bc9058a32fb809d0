"""The reference's train.py, unmodified, binds to the B200 implementation through ideas_b200/compat
(SURVEY.md §8b "Python surface train.py binds to").  Needs /root/reference, so it runs in the build
container only (skipped on the GPU box); nothing is executed on a device."""
import importlib
import os
import sys

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train.py")), reason="reference tree not present")
def test_reference_train_imports_our_modules():
    from ideas_b200.compat import run_train
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    undo = lambda: None  # noqa: E731
    try:
        undo = run_train.install_shims()
        sys.path.append(REF)                      # only `train` itself resolves from the reference
        for name in ("train",):
            sys.modules.pop(name, None)
        train = importlib.import_module("train")
        import ideas_b200.models as M
        import ideas_b200.utils as U
        assert train.__file__.startswith(REF)
        assert train.init_model is M.init_model
        assert train.patchify_image is U.patchify_image and train.d_r1_loss is U.d_r1_loss
        CU = sys.modules["utils"]                  # the compat alias module (returns the message on the CPU)
        assert CU.__file__.endswith(os.path.join("compat", "utils.py"))
        assert train.tensor_to_message is CU.tensor_to_message and train.accumulate is U.accumulate
        ds = train.set_dataset(type="normal", path="synthetic:8", transform=None, resolution=32)
        assert len(ds) == 8
        assert tuple(ds[0].shape) == (3, 32, 32) and float(ds[0].min()) >= -1.0
        import torch
        opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-3, betas=(0, 0.99))   # train.py:417-426
        assert opt.defaults["betas"] == (0.0, 0.99)
    finally:
        undo()
        sys.path[:] = saved_path
        for k in list(sys.modules):
            # drop only the reference-facing aliases; torch sub-modules register dispatcher libraries at
            # import and must never be imported twice in one process
            if k not in saved_mods and k.split(".")[0] in ("train", "models", "utils", "dataset", "stylegan2"):
                del sys.modules[k]


def test_time_change_format():
    from ideas_b200.utils import time_change
    assert time_change(5.9) == "5s" and time_change(125) == "2m 5s" and time_change(3725) == "1h 2m 5s"
    assert time_change(3600) == "60m 0s" and time_change(60) == "60s"       # the reference's strict ">" boundaries
