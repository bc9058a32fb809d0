"""ctypes binding of libideas_b200.so (the C ABI declared in include/ideas_b200.h).

The shared library is built in-tree by ``build()`` (nvcc, sm_100a only) and loaded once.
There is NO fallback: if the library is missing and cannot be built, or an entry point
reports an error, a RuntimeError is raised -- mirroring the reference, whose ops raise
from TORCH_CHECK (stylegan2/op/fused_bias_act.cpp:7-14).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libideas_b200.so")
SOURCES = ["bias_act.cu", "upfirdn2d.cu", "conv_simt.cu", "conv_umma.cu", "conv_pointwise.cu", "elementwise.cu", "bits.cu",
           "linear.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None

IMPL_AUTO, IMPL_SIMT, IMPL_UMMA = 0, 1, 2
ACT_NONE, ACT_LRELU = 0, 1

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_P = c_void_p
_PROTOTYPES = {
    "ideas_abi_version": [],
    "ideas_device_cc": [],
    "ideas_umma_available": [],
    "ideas_set_option": [ctypes.c_char_p, c_int],
    "ideas_fused_bias_act": [_P, _P, _P, _P, c_int, c_int, c_float, c_float, c_int64, c_int64, c_int, _P],
    "ideas_bias_act_backward": [_P, _P, _P, _P, c_float, c_float, c_int64, c_int64, c_int, _P],
    "ideas_modconv_act_backward": [_P, _P, _P, _P, _P, _P, c_float, c_float, c_int, c_int64, c_int, _P],
    "ideas_upfirdn2d": [_P, _P, _P] + [c_int] * 14 + [_P, c_float, c_float, _P],
    "ideas_upfirdn2d_res": [_P, _P, _P] + [c_int] * 14 + [_P, c_float, c_float, _P, c_float, _P],
    "ideas_blur_act_backward": [_P, _P, _P, _P, _P] + [c_int] * 10 + [c_float, c_float, _P],
    "ideas_blur_scale_dot_backward": [_P, _P, _P, _P, _P, _P] + [c_int] * 10 + [_P],
    "ideas_blur_bias_act_post": [_P, _P, _P, _P, _P, _P] + [c_int] * 10 + [c_float, c_float, _P],
    "ideas_pack_weight": [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P],
    "ideas_unpack_weight_grad": [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P],
    "ideas_repack_dgrad": [_P, _P, c_int, c_int, c_int, _P],
    "ideas_conv2d_forward": [_P] * 6 + [c_int] * 9 + [c_int, c_float, c_float, c_int, _P],
    "ideas_conv2d_forward_res": [_P] * 7 + [c_float] + [c_int] * 9 + [c_int, c_float, c_float, c_int, _P],
    "ideas_conv2d_dgrad": [_P] * 6 + [c_int] * 11 + [c_int, c_float, c_float, c_int, _P],
    "ideas_conv2d_wgrad": [_P] * 5 + [c_int] * 11 + [c_int, _P],
    "ideas_scale_channels": [_P, _P, _P, c_int, c_int64, c_int, _P],
    "ideas_channel_dot": [_P, _P, _P, _P, _P, c_int, c_int64, c_int, _P],
    "ideas_add_scale": [_P, _P, _P, c_float, c_int64, _P],
    "ideas_gemm_nt": [_P, _P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_int, _P],
    "ideas_reflect_pad2d": [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "ideas_patchify_forward": [_P, _P, _P] + [c_int] * 7 + [_P],
    "ideas_patchify_backward": [_P, _P, _P] + [c_int] * 7 + [_P],
    "ideas_bits_encode": [_P, _P, _P, c_int, c_int, c_int, c_float, _P],
    "ideas_bits_decode": [_P, _P, c_int, c_int, c_int, _P],
    "ideas_bits_count_errors": [_P, _P, _P, c_int64, _P],
}
EXPORTS = sorted(list(_PROTOTYPES) + ["ideas_last_error", "ideas_launch_count"])


def launch_count() -> int:
    """Kernels launched by the library so far in this process."""
    return int(lib().ideas_launch_count())


def _source_digest() -> str:
    """sha256 over every CUDA source, header and the compile flags: the library is stale exactly when this changes
    (modification times do not survive the snapshot that carries the tree to a GPU box)."""
    import hashlib
    deps = sorted([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))])
    deps.append(os.path.join(_HERE, "..", "include", "ideas_b200.h"))
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    for d in deps:
        with open(d, "rb") as f:
            h.update(os.path.basename(d).encode() + b"\0" + f.read())
    return h.hexdigest()


def _is_current(digest: str) -> bool:
    try:
        with open(LIB_PATH + ".srchash") as f:
            return os.path.exists(LIB_PATH) and f.read().strip() == digest
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ideas_b200/lib/libideas_b200.so."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    digest = _source_digest()
    if not force and _is_current(digest):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    objdir = os.path.join(_HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared", "--use_fast_math=false")]
    for s in srcs:  # one nvcc per translation unit, in parallel
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc] + flags + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose and out.strip():
            print(out)
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    os.replace(tmp, LIB_PATH)
    with open(LIB_PATH + ".srchash", "w") as f:
        f.write(digest)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (building if necessary) the shared library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        build()                              # no-op when the library matches the sources (content digest)
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _PROTOTYPES.items():
            fn = getattr(L, name)           # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = c_int
        L.ideas_last_error.argtypes = []
        L.ideas_last_error.restype = ctypes.c_char_p
        L.ideas_launch_count.argtypes = []
        L.ideas_launch_count.restype = ctypes.c_ulonglong
        _lib = L
        # kernel-selection options for A/B measurements, e.g. IDEAS_OPTS=pmh=0,halo=2 (see ideas_set_option)
        for kv in filter(None, os.environ.get("IDEAS_OPTS", "").split(",")):
            name, val = kv.split("=")
            check(L.ideas_set_option(name.encode(), int(val)), "ideas_set_option")
    return _lib


def umma_enabled() -> bool:
    """True when the tcgen05/TMA convolution kernels are compiled in and not disabled by
    IDEAS_B200_UMMA=0 (the SIMT fp32 path then serves every shape)."""
    if os.environ.get("IDEAS_B200_UMMA", "1") == "0":
        return False
    L = lib()
    return bool(L.ideas_umma_available())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ideas_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libideas_b200: {what} failed (code {rc}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)
