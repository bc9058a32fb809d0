"""ideas_b200 -- B200-native implementation of the IDEAS GAN-conv hot path.

Python host side above the C ABI (include/ideas_b200.h):
  ideas_b200.stylegan2.op     FusedLeakyReLU / fused_leaky_relu / upfirdn2d (+ conv family)
  ideas_b200.stylegan2.model  ModulatedConv2d, StyledConv*, ToRGB, EqualConv2d, EqualLinear, Blur ...
  ideas_b200.models           the seven IDEAS networks + init_model
  ideas_b200.utils            losses, R1, patchify, EMA, bit <-> tensor mapping (integer kernels)
  ideas_b200.train_step       one IDEAS training iteration (+ data-parallel gradient all-reduce)
"""
from . import _lib

__all__ = ["_lib"]
__version__ = "0.1.0"
