"""One IDEAS training iteration over the functional oracle nets (TEST INFRASTRUCTURE).

Restates the loop body train.py:33-221 of the reference (D phase, lazy R1, G/E/Ex
phase with its two backwards, EMA) plus the helpers it calls in utils.py:42-66,105-149.
Used (a) as the CPU baseline timed by bench.py and (b) by tests to compare losses and
gradients of the CUDA training step on identical inputs; all randomness can be supplied
through ``draws`` so both sides see the same Z, T2 and crop boxes.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import nets

NETS = {"E": "DisentanglementEncoder", "G": "Generator", "Gstru": "StructureGenerator",
        "Ex": "TensorExtractor", "Dreal": "ImageLevelDiscriminator",
        "Dco": "CooccurenceDiscriminator", "Ddist": "DistributionDiscriminator"}
EMA_NETS = ("E", "G", "Gstru", "Ex")


# ---- utils.py:105-124 -------------------------------------------------------------
def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    (g,) = torch.autograd.grad(real_pred.sum(), real_img, create_graph=True)
    return g.pow(2).reshape(g.shape[0], -1).sum(1).mean()


# ---- utils.py:127-149 -------------------------------------------------------------
def draw_crops(n_crop, height, width, min_size=1 / 8, max_size=1 / 4):
    """The random part of patchify_image: crop sizes from torch.rand, corners from
    random.randrange (utils.py:128-138).  Returns [(y, x, h, w)]."""
    size = torch.rand(n_crop) * (max_size - min_size) + min_size
    hs = (size * height).type(torch.int64).tolist()
    ws = (size * width).type(torch.int64).tolist()
    return [(random.randrange(0, height - h), random.randrange(0, width - w), h, w) for h, w in zip(hs, ws)]


def patchify_image(img, crops, max_size=1 / 4):
    B, C, H, W = img.shape
    th, tw = int(H * max_size), int(W * max_size)
    out = [F.interpolate(img[:, :, y:y + h, x:x + w], size=(th, tw), mode="bilinear", align_corners=False)
           for (y, x, h, w) in crops]
    return torch.stack(out, 1).reshape(-1, C, th, tw)


class OracleTrainer:
    """Holds the 7 networks (+4 EMA copies) as state_dicts of leaf tensors and the three
    Adam optimisers of train.py:417-432."""

    def __init__(self, *, channel=32, structure_channel=8, texture_channel=2048, N=1,
                 image_size=256, channel_multiplier=1, lr=0.002, d_reg_every=16,
                 n_crop=8, ref_crop=4, lambda_Ex=10.0, real_r1=10.0, texture_r1=1.0,
                 dist_r1=1.0, num_iters=80000, seed: Optional[int] = 0, states=None):
        self.cfg = dict(channel=channel, structure_channel=structure_channel,
                        texture_channel=texture_channel, N=N, image_size=image_size,
                        channel_multiplier=channel_multiplier)
        self.N, self.d_reg_every, self.n_crop, self.ref_crop = N, d_reg_every, n_crop, ref_crop
        self.lambda_Ex, self.real_r1, self.texture_r1, self.dist_r1 = lambda_Ex, real_r1, texture_r1, dist_r1
        self.num_iters = num_iters
        if seed is not None:
            torch.manual_seed(seed)
        self.sd: Dict[str, Dict[str, torch.Tensor]] = {}
        order = ["E", "G", "Gstru", "Ex", "Dreal", "Dco", "Ddist"]      # train.py:390-399
        for k in order:
            st = states[k] if states is not None else nets.init_state(NETS[k], **self.cfg)
            self.sd[k] = {n: t.detach().clone() for n, t in st.items()}
        self.ema = {}
        for k in EMA_NETS:                                              # train.py:400-414 (decay 0 => copy)
            if states is None:
                nets.init_state(NETS[k], **self.cfg)                    # consume the same RNG draws
            self.ema[k] = {n: t.detach().clone() for n, t in self.sd[k].items()}
        ratio = d_reg_every / (d_reg_every + 1)
        self.g_optim = torch.optim.Adam(self.params("E") + self.params("G") + self.params("Gstru"),
                                        lr=lr, betas=(0.0, 0.99))
        self.ex_optim = torch.optim.Adam(self.params("Ex"), lr=lr, betas=(0.0, 0.99))
        self.d_optim = torch.optim.Adam(self.params("Dreal") + self.params("Dco") + self.params("Ddist"),
                                        lr=lr * ratio, betas=(0.0 ** ratio, 0.99 ** ratio))
        self.accum = 0.5 ** (32 / (10 * 1000))                          # train.py:30

    def params(self, k) -> List[torch.Tensor]:
        return [t for n, t in self.sd[k].items() if not nets.is_buffer(n)]

    def requires_grad(self, k, flag):                                   # utils.py:50-52
        for t in self.params(k):
            t.requires_grad_(flag)

    def net(self, k, *a, **kw):
        return nets.FORWARD[NETS[k]](self.sd[k], *a, **kw)

    def accumulate(self, k):                                            # utils.py:55-60
        with torch.no_grad():
            for n, t in self.sd[k].items():
                if not nets.is_buffer(n):
                    self.ema[k][n].mul_(self.accum).add_(t, alpha=1 - self.accum)

    def draw(self, X, S_shape, T_shape):
        """All randomness one phase needs, in the order train.py draws it."""
        H, W = X.shape[2], X.shape[3]
        d = {"Z": torch.rand(S_shape[0], self.N, S_shape[2], S_shape[3]) * 2 - 1,
             "T2": torch.rand(T_shape) * 2 - 1}
        d["fake_crops"] = draw_crops(self.n_crop, H, W)
        return d

    # ------------------------------------------------------------------ train.py:33-221
    def step(self, X: torch.Tensor, iter_idx: int, draws: Optional[dict] = None) -> Dict[str, float]:
        H, W = X.shape[2], X.shape[3]
        draws = dict(draws or {})
        losses = {}
        # ---- D phase (train.py:48-102)
        for k in EMA_NETS:
            self.requires_grad(k, False)
        for k in ("Dreal", "Dco", "Ddist"):
            self.requires_grad(k, True)
        S1, T1 = self.net("E", X)
        Z = draws.get("Z_d")
        if Z is None:
            Z = torch.rand(S1.shape[0], self.N, S1.shape[2], S1.shape[3]) * 2 - 1
        S2 = self.net("Gstru", Z)
        T2 = draws.get("T2_d")
        if T2 is None:
            T2 = torch.rand_like(T1) * 2 - 1
        hat_X1, hat_X2, hat_X3 = self.net("G", S1, T1), self.net("G", S2, T1), self.net("G", S2, T2)
        fake_pred = self.net("Dreal", torch.cat((hat_X1, hat_X2, hat_X3), 0))
        real_pred = self.net("Dreal", X)
        D_real = d_logistic_loss(real_pred, fake_pred)
        fc = draws.get("fake_crops_d") or draw_crops(self.n_crop, H, W)
        rc = draws.get("real_crops_d") or draw_crops(self.n_crop, H, W)
        fc_ref = draws.get("ref_crops_d") or draw_crops(self.ref_crop * self.n_crop, H, W)
        fake_patch, real_patch = patchify_image(hat_X2, fc), patchify_image(X, rc)
        ref_patch = patchify_image(X, fc_ref)
        fake_tex, ref_input = self.net("Dco", fake_patch, ref_patch, ref_batch=self.ref_crop)
        real_tex, _ = self.net("Dco", real_patch, ref_input=ref_input)
        D_tex = d_logistic_loss(real_tex, fake_tex)
        D_dist = d_logistic_loss(self.net("Ddist", T2), self.net("Ddist", T1))
        self.d_optim.zero_grad()
        (D_real + D_tex + D_dist).backward()
        self.d_optim.step()
        losses.update(D_real_loss=D_real.item(), D_texture_loss=D_tex.item(), D_dist_loss=D_dist.item())
        # ---- lazy R1 (train.py:105-129)
        if iter_idx % self.d_reg_every == 0:
            Xr = X.detach().clone().requires_grad_(True)
            r1_real = d_r1_loss(self.net("Dreal", Xr), Xr)
            rp = real_patch.detach().clone().requires_grad_(True)
            pred, _ = self.net("Dco", rp, ref_patch, ref_batch=self.ref_crop)
            r1_tex = d_r1_loss(pred, rp)
            T2r = T2.detach().clone().requires_grad_(True)
            r1_dist = d_r1_loss(self.net("Ddist", T2r), T2r)
            self.d_optim.zero_grad()
            tot = self.real_r1 / 3 * r1_real * self.d_reg_every
            tot = tot + self.texture_r1 / 3 * r1_tex * self.d_reg_every
            tot = tot + self.dist_r1 / 3 * r1_dist * self.d_reg_every
            tot.backward()
            self.d_optim.step()
            losses.update(D_real_r1_loss=r1_real.item(), D_texture_r1_loss=r1_tex.item(),
                          D_dist_r1_loss=r1_dist.item())
        # ---- G / E / Ex phase (train.py:135-216)
        for k in EMA_NETS:
            self.requires_grad(k, True)
        for k in ("Dreal", "Dco", "Ddist"):
            self.requires_grad(k, False)
        S1, T1 = self.net("E", X)
        Z = draws.get("Z_g")
        if Z is None:
            Z = torch.rand(S1.shape[0], self.N, S1.shape[2], S1.shape[3]) * 2 - 1
        S2 = self.net("Gstru", Z)
        T2 = draws.get("T2_g")
        if T2 is None:
            T2 = torch.rand_like(T1) * 2 - 1
        hat_X1, hat_X2, hat_X3 = self.net("G", S1, T1), self.net("G", S2, T1), self.net("G", S2, T2)
        G_rec = F.l1_loss(hat_X1, X)
        G_real = g_nonsaturating_loss(self.net("Dreal", torch.cat((hat_X1, hat_X2, hat_X3), 0)))
        E_dist = g_nonsaturating_loss(self.net("Ddist", T1))
        fc = draws.get("fake_crops_g") or draw_crops(self.n_crop, H, W)
        fc_ref = draws.get("ref_crops_g") or draw_crops(self.ref_crop * self.n_crop, H, W)
        pred, _ = self.net("Dco", patchify_image(hat_X2, fc), patchify_image(X, fc_ref), ref_batch=self.ref_crop)
        G_tex = g_nonsaturating_loss(pred)
        container = hat_X3 if iter_idx > self.num_iters * 0.8 else hat_X2
        hat_S2, _ = self.net("E", container)
        E_stru = F.l1_loss(hat_S2, S2)
        Ex_loss = F.l1_loss(self.net("Ex", hat_S2), Z)
        total = (G_rec + G_tex + 2 * G_real) + (E_dist + E_stru) + self.lambda_Ex * Ex_loss
        self.g_optim.zero_grad()
        total.backward(retain_graph=True)
        self.g_optim.step()
        self.ex_optim.zero_grad()
        Ex_loss.backward()
        self.ex_optim.step()
        for k in EMA_NETS:
            self.accumulate(k)
        losses.update(G_rec_loss=G_rec.item(), G_real_loss=G_real.item(), G_texture_loss=G_tex.item(),
                      E_dist_loss=E_dist.item(), E_stru_loss=E_stru.item(), Ex_loss=Ex_loss.item(),
                      Loss_total=total.item())
        return losses
