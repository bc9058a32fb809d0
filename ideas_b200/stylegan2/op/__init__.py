from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d, blur_bias_act
from .conv import conv2d, conv_transpose2d

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "blur_bias_act", "conv2d", "conv_transpose2d"]
