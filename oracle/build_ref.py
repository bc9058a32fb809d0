"""Compile the reference's own two CUDA operators for sm_100a into oracle/_ref/ (TEST / BENCH INFRASTRUCTURE).

    python oracle/build_ref.py            # needs /root/reference (build container); no-op message otherwise

Sources are compiled where they lie -- /root/reference/stylegan2/op/{fused_bias_act.cpp, fused_bias_act_kernel.cu,
upfirdn2d.cpp, upfirdn2d_kernel.cu} -- with torch.utils.cpp_extension (the reference builds them the same way at
import, fused_act.py:11-17, upfirdn2d.py:10-16); nothing is copied into the repository, only the two extension
modules land in oracle/_ref/ (git-ignored, shipped to the GPU box with the working tree):

    oracle/_ref/ideas_ref_fused.so      fused_bias_act(input, bias, refer, act, grad, alpha, scale)   "K1"
    oracle/_ref/ideas_ref_upfirdn2d.so  upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pads...)  "K2/K3"

They are the "reference op/" comparator of BASELINE.json configs[1] (bench.py `library_baseline`) and a second
checker for the bias-act / upfirdn2d parity tests on the GPU.  ``load()`` returns (fused, upfirdn2d) modules or None.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_OP = os.path.join(os.environ.get("IDEAS_REFERENCE", "/root/reference"), "stylegan2", "op")
MODULES = {"ideas_ref_fused": ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"],
           "ideas_ref_upfirdn2d": ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"]}


def build(verbose: bool = False) -> bool:
    """Returns True when both modules exist afterwards."""
    if all(os.path.exists(os.path.join(OUT, n + ".so")) for n in MODULES):
        return True
    if not os.path.isdir(REF_OP):
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension
    for name, files in MODULES.items():
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        cpp_extension.load(name, sources=[os.path.join(REF_OP, f) for f in files], build_directory=bdir,
                           extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
                           verbose=verbose, is_python_module=True)
        os.replace(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
    return True


def load():
    import torch  # noqa: F401  (the extension modules link against libtorch)
    mods = []
    for name in MODULES:
        path = os.path.join(OUT, name + ".so")
        if not os.path.exists(path):
            return None
        if name in sys.modules:
            mods.append(sys.modules[name])
            continue
        spec = importlib.util.spec_from_loader(name, importlib.machinery.ExtensionFileLoader(name, path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[name] = mod
        mods.append(mod)
    return tuple(mods)


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference tree not present, nothing built")
