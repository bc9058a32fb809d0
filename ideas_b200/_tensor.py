"""Small host-side helpers: NHWC layout handling, raw pointers and the current stream.

Activations live in HBM as NHWC (channels-last): a logical (N, C, H, W) torch tensor whose
memory is (N, H, W, C)-dense.  Module inputs in any other layout are converted once at the
boundary; every op in this package returns channels-last tensors.
"""
from __future__ import annotations

import ctypes

import torch


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            # same behaviour as the reference's CHECK_CUDA (stylegan2/op/upfirdn2d.cpp:8-16)
            raise RuntimeError("ideas_b200 ops need CUDA tensors (there is no CPU fallback); got device %s" % t.device)
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("ideas_b200 ops compute in fp32; got %s" % t.dtype)


def is_nhwc_dense(x: torch.Tensor) -> bool:
    return x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous()


def nhwc(x: torch.Tensor) -> torch.Tensor:
    """Return x (logical NCHW) with NHWC-dense memory, copying only if needed."""
    if is_nhwc_dense(x):
        return x
    return x.contiguous(memory_format=torch.channels_last)


def empty_nhwc(n: int, c: int, h: int, w: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty((n, h, w, c), device=like.device, dtype=like.dtype).permute(0, 3, 1, 2)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
