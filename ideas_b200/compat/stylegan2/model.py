"""`stylegan2.model` (reference models.py:7 imports from it) -> the B200 layer library."""
from ideas_b200.stylegan2.model import *  # noqa: F401,F403
from ideas_b200.stylegan2.model import (Blur, EqualConv2d, EqualLinear, ModulatedConv2d, ScaledLeakyReLU,  # noqa: F401
                                        StyledConv, StyledConv_without_noise, ToRGB, make_kernel)
