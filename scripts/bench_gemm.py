"""EqualLinear-shaped GEMMs (a @ b^T, skinny M = batch) through ideas_gemm_nt: reduction split rule old (gemm_split=0)
vs CTAs-per-SM targets."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200.stylegan2.op.linear import matmul_nt

dev = torch.device("cuda")
# M, N, R
CASES = [(32, 5120, 2048), (32, 2048, 2048), (32, 512, 512), (96, 512, 8192), (32, 512, 8192), (32, 256, 256), (2048, 32, 5120),
         (5120, 2048, 32), (32, 384, 2048)]
TARGETS = (0, 2, 4, 8, 16)
print(f"{'M':>5s} {'N':>5s} {'R':>5s} | " + " | ".join(f"split={t:<2d} us" for t in TARGETS))
for M, N, R in CASES:
    a = torch.randn(M, R, device=dev)
    b = torch.randn(N, R, device=dev)
    cells = []
    for t in TARGETS:
        _lib.call("ideas_set_option", b"gemm_split", t)
        for _ in range(3):
            matmul_nt(a, b)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            matmul_nt(a, b)
        e.record()
        torch.cuda.synchronize()
        cells.append(s.elapsed_time(e) / 20 * 1e3)
    print(f"{M:5d} {N:5d} {R:5d} | " + " | ".join(f"{c:11.1f}" for c in cells), flush=True)
_lib.call("ideas_set_option", b"gemm_split", 16)
