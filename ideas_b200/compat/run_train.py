"""Launcher: run the reference's train.py UNMODIFIED on the B200 implementation.

    python -m ideas_b200.compat.run_train /path/to/IDEAS/train.py --exp_name demo \
        --dataset_type normal --dataset_path /data/ffhq --num_iters 80000 --batch_size 32 ...

``--dataset_path synthetic[:length]`` (with either dataset type) selects LSUN-shaped random tensors instead of
files (ideas_b200/compat/dataset.py); ``lmdb`` needs the `lmdb` package.

Shims applied outside the reference tree (SURVEY.md App. D -- the reference no longer runs as written
on torch >= 2.6 / torchvision >= 0.13):
  * `models`, `utils`, `dataset`, `stylegan2` resolve to ideas_b200/compat (this directory first on sys.path);
  * torch.optim.Adam: integer betas (train.py:417-426 passes `0`) are coerced to float;
  * torchvision.utils.save_image / make_grid: the removed `range=` keyword maps to `value_range=`;
  * torch.load: `weights_only=False` so checkpoints holding the argparse Namespace load (train.py:437).
"""
import os
import runpy
import sys


def install_shims():
    """Apply the shims; returns a callable that undoes the monkey-patches (used by the tests)."""
    import torch
    undo = []
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    for name in ("models", "utils", "dataset", "stylegan2", "stylegan2.model", "stylegan2.op"):
        sys.modules.pop(name, None)

    adam_init = torch.optim.Adam.__init__

    def patched_adam(self, params, *a, **kw):
        if "betas" in kw:
            kw["betas"] = tuple(float(b) for b in kw["betas"])
        return adam_init(self, params, *a, **kw)

    torch.optim.Adam.__init__ = patched_adam
    undo.append(lambda: setattr(torch.optim.Adam, "__init__", adam_init))
    try:
        import torchvision.utils as tvu
        for fname in ("save_image", "make_grid"):
            fn = getattr(tvu, fname)

            def wrap(*a, _fn=fn, **kw):
                if "range" in kw:
                    kw["value_range"] = kw.pop("range")
                return _fn(*a, **kw)

            setattr(tvu, fname, wrap)
            undo.append(lambda _n=fname, _f=fn: setattr(tvu, _n, _f))
    except Exception:  # torchvision is optional for the benchmark path
        pass
    load = torch.load

    def patched_load(*a, **kw):
        kw.setdefault("weights_only", False)
        return load(*a, **kw)

    torch.load = patched_load
    undo.append(lambda: setattr(torch, "load", load))
    return lambda: [u() for u in undo]


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script, rest = argv[0], argv[1:]
    install_shims()
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
