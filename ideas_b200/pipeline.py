"""Sender / receiver pipeline of IDEAS on the B200 networks (reference train.py:249-293).

    sender:    message bits -> secret tensor Z (utils.message_to_tensor) -> structure code S2 = Gstru(Z)
               -> container image = G(S2, T) with a texture vector T
    receiver:  image -> S = E(image)[0] -> Z^ = Ex(S) -> bits (utils.tensor_to_message), BER by XOR/popcount

The networks are the (EMA) modules of a ``Trainer`` or any dict with the keys ``Gstru``, ``G``, ``E``, ``Ex``.
Messages travel bit-packed (32 per int32 word) on the device; nothing runs on the CPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import utils as U


def _net(nets, key):
    return nets[key + "_ema"] if (key + "_ema") in nets else nets[key]


@torch.no_grad()
def hide(nets: Dict[str, torch.nn.Module], message: torch.Tensor, texture: torch.Tensor, sigma: int = 1,
         delta: float = 0.5, N: int = 1, image_size: int = 256) -> torch.Tensor:
    """message (B, sigma*N*h*w) of 0/1 with h = w = image_size/16, texture (B, texture_channel) in [-1, 1]
    -> container images (B, 3, image_size, image_size).  Mirrors train.py:254-262."""
    hw = image_size // 16
    if message.shape[1] != sigma * N * hw * hw:
        raise ValueError(f"message must hold sigma*N*h*w = {sigma * N * hw * hw} bits per image, got {message.shape[1]}")
    dev = texture.device
    z = U.message_to_tensor(message, sigma, delta, device=dev).reshape(message.shape[0], N, hw, hw)
    return _net(nets, "G")(_net(nets, "Gstru")(z), texture)


@torch.no_grad()
def extract(nets: Dict[str, torch.nn.Module], image: torch.Tensor, sigma: int = 1) -> torch.Tensor:
    """container images -> recovered message (B, sigma*N*h*w) of 0/1 floats on the device (train.py:272-279)."""
    s, _ = _net(nets, "E")(image)
    z_hat = _net(nets, "Ex")(s)
    return U.tensor_to_message(z_hat.reshape(z_hat.shape[0], -1), sigma)


@torch.no_grad()
def round_trip_ber(nets: Dict[str, torch.nn.Module], message: torch.Tensor, texture: torch.Tensor, sigma: int = 1,
                   delta: float = 0.5, N: int = 1, image_size: int = 256,
                   attack: Optional[callable] = None) -> float:
    """BER = mean |M - M^| of hide -> (optional channel ``attack``) -> extract (train.py:285), counted on the
    device with XOR + popcount over the packed words."""
    img = hide(nets, message, texture, sigma, delta, N, image_size)
    if attack is not None:
        img = attack(img)
    got = extract(nets, img, sigma)
    a = U.pack_message(message.to(got.device))
    b = U.pack_message(got)
    return U.bit_error_rate(a, b, message.numel())
