"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads, and
exports every symbol include/ideas_b200.h declares; argument validation (no GPU needed: the
checks run before any launch); the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ideas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ideas_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from ideas_b200 import _lib
    path = _lib.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ideas_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(names), set(_lib.EXPORTS) - set(names)
    assert _lib.lib().ideas_abi_version() >= 1


def test_sass_is_sm100a_only():
    from ideas_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _lib.build()], capture_output=True, text=True)
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_without_a_gpu():
    from ideas_b200 import _lib
    L = _lib.lib()
    p0 = ctypes.c_void_p(0)
    # negative element count
    assert L.ideas_fused_bias_act(p0, p0, p0, p0, 3, 0, 0.2, 1.0, -1, 1, 1, p0) == -1
    assert b"negative" in L.ideas_last_error()
    # empty FIR kernel / bad factors
    assert L.ideas_upfirdn2d(p0, p0, p0, 1, 4, 4, 4, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, p0, 0.2, 1.0, p0) == -1
    assert L.ideas_upfirdn2d(p0, p0, p0, 1, 4, 4, 4, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, p0, 0.2, 1.0, p0) == -1
    # conv: kernel with more than 16 taps, kernel larger than the padded input
    assert L.ideas_conv2d_forward(p0, p0, p0, p0, p0, p0, 1, 8, 8, 4, 4, 5, 5, 1, 0, 0, 0.2, 1.0, 0, p0) == -1
    assert L.ideas_conv2d_forward(p0, p0, p0, p0, p0, p0, 1, 2, 2, 4, 4, 3, 3, 1, 0, 0, 0.2, 1.0, 0, p0) == -1
    # empty problems succeed without touching the device
    assert L.ideas_fused_bias_act(p0, p0, p0, p0, 3, 0, 0.2, 1.0, 0, 1, 1, p0) == 0
    assert L.ideas_conv2d_forward(p0, p0, p0, p0, p0, p0, 0, 8, 8, 4, 4, 3, 3, 1, 1, 0, 0.2, 1.0, 0, p0) == 0
    assert L.ideas_bits_decode(p0, p0, 0, 16, 1, p0) == 0
    # the two entry points added for the fused backward / NHWC reflection pad
    assert L.ideas_blur_act_backward(p0, p0, p0, p0, p0, 1, 0, 4, 4, 4, 4, 1, 1, 1, 1, 0.2, 1.0, p0) == -1      # in_h = 0
    assert L.ideas_blur_act_backward(p0, p0, p0, p0, p0, 0, 4, 4, 4, 4, 4, 1, 1, 1, 1, 0.2, 1.0, p0) == 0       # empty batch
    assert L.ideas_reflect_pad2d(p0, p0, 1, 4, 4, 4, 4, 0, p0) == -1                                            # pad >= size
    assert b"padding" in L.ideas_last_error()
    assert L.ideas_reflect_pad2d(p0, p0, 0, 4, 4, 4, 1, 0, p0) == 0
    # tuning knobs: known names accepted, unknown rejected with a message
    assert L.ideas_set_option(b"halo", 1) == 0 and L.ideas_set_option(b"wgrad_reuse", 1) == 0
    # the knobs INTEGRATION.md lists, set to their defaults
    for name, default in ((b"pmh", 1), (b"pmh_resident", 1), (b"tail_split", 1), (b"tma_tf32", 1), (b"blur_variant", 0),
                          (b"pair", 1), (b"pair_epi", 1), (b"pair_stages", 0), (b"dgrad_phases", 1), (b"halo_epi", 1),
                          (b"resample_variant", 0), (b"gemm_split", 16)):
        assert L.ideas_set_option(name, default) == 0, name
    assert L.ideas_set_option(b"dgrad_phases", 7) == -1          # out of range
    assert L.ideas_set_option(b"dgrad_phases", 1) == 0
    assert L.ideas_set_option(b"no_such_option", 1) == -1
    assert b"unknown option" in L.ideas_last_error()


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "ideas_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ideas_b200.stylegan2.op import fused_leaky_relu, upfirdn2d
    from ideas_b200.stylegan2.model import make_kernel
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fused_leaky_relu(torch.randn(1, 4, 2, 2), torch.zeros(4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        upfirdn2d(torch.randn(1, 4, 8, 8), make_kernel([1, 3, 3, 1]), pad=(2, 1))


def test_pipeline_validates_message_length():
    """ideas_b200.pipeline.hide checks the message geometry before touching the device (train.py:254-257)."""
    import torch
    from ideas_b200 import pipeline
    with pytest.raises(ValueError, match="sigma\\*N\\*h\\*w"):
        pipeline.hide({}, torch.zeros(2, 100), torch.zeros(2, 2048), sigma=1, N=1, image_size=256)
