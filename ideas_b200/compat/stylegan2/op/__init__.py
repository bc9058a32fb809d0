"""`stylegan2.op` (reference stylegan2/op/__init__.py:1-2) -> the B200 ops."""
from ideas_b200.stylegan2.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401
