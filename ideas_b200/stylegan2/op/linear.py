"""EqualLinear's GEMM on the hand-written fp32 kernel (A6; reference: F.linear -> cuBLAS, stylegan2/model.py:151-161).

``matmul_nt(a, b)`` = a @ b.T for 2-D CUDA tensors of any strides.  Its backward is written with itself on
transposed views, so gradients of every order exist (the R1 penalty differentiates the discriminators' linear heads
twice, utils.py:112-118)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from ... import _lib
from ..._tensor import ptr, require_cuda, stream_ptr


def _gemm_nt(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    m, r = a.shape
    n, r2 = b.shape
    if r != r2:
        raise RuntimeError(f"matmul_nt: inner dimensions differ: {tuple(a.shape)} x {tuple(b.shape)}^T")
    c = torch.zeros((m, n), device=a.device, dtype=a.dtype)
    _lib.call("ideas_gemm_nt", ptr(c), ptr(a), ptr(b), m, n, r, a.stride(0), a.stride(1), b.stride(0), b.stride(1), n,
              float(alpha), 1, stream_ptr(a))
    return c


class MatMulNT(Function):
    @staticmethod
    def forward(ctx, a, b):
        require_cuda(a, b)
        ctx.save_for_backward(a, b)
        return _gemm_nt(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = MatMulNT.apply(g, b.t()) if ctx.needs_input_grad[0] else None           # (M,N) x (R,N)^T -> (M,R)
        gb = MatMulNT.apply(g.t(), a.t()) if ctx.needs_input_grad[1] else None       # (N,M) x (R,M)^T -> (N,R)
        return ga, gb


def matmul_nt(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a (M, R) @ b (N, R)^T -> (M, N), exact fp32."""
    return MatMulNT.apply(a, b)
