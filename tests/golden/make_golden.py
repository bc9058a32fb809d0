#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the
GPU box):

    TORCH_EXTENSIONS_DIR=/tmp/torch_ext TORCH_CUDA_ARCH_LIST=10.0a \
        python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so parity is
pinned to outputs of the reference code itself, executed here on CPU with fixed seeds:
its own CPU op paths (fused_act.py:87-94, upfirdn2d.py:159-200), its nn.Modules
(stylegan2/model.py, models.py), its bit mapping (utils.py:74-97) and its training loop
(train.py:21-221, driven through the compatibility shims of SURVEY.md App. D -- nothing
in the reference tree is edited).  Everything is saved as small .pt / .json files.
"""
import argparse
import hashlib
import json
import os
import random
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def sha_state(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def ns(**kw):
    d = dict(channel=32, structure_channel=8, texture_channel=2048, N=1, image_size=256,
             channel_multiplier=1, blur_kernel=(1, 3, 3, 1))
    d.update(kw)
    return argparse.Namespace(**d)


NET_NAMES = ["DisentanglementEncoder", "Generator", "StructureGenerator", "TensorExtractor",
             "ImageLevelDiscriminator", "CooccurenceDiscriminator", "DistributionDiscriminator"]


def main():
    sys.path.insert(0, REF)
    import models as R                      # noqa: E402  (JIT-builds the reference CUDA ops once)
    import utils as RU                      # noqa: E402
    from stylegan2 import model as RM       # noqa: E402
    from stylegan2.op import fused_leaky_relu as ref_flrelu  # noqa: E402
    from stylegan2.op.upfirdn2d import upfirdn2d_native      # noqa: E402

    torch.set_num_threads(8)

    # ------------------------------------------------------------------ 1. op-level goldens
    ops = {}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 5, 7, 9, generator=g)
    b = torch.randn(5, generator=g)
    ops["flrelu"] = dict(x=x, bias=b, out=ref_flrelu(x, b))
    x2 = torch.randn(3, 6, generator=g)
    b2 = torch.randn(6, generator=g)
    ops["flrelu_2d"] = dict(x=x2, bias=b2, out=ref_flrelu(x2, b2))

    k4 = RM.make_kernel([1, 3, 3, 1])
    k3 = RM.make_kernel([1, 2, 1])
    k2 = RM.make_kernel([1, 1])
    k12 = RM.make_kernel([float(i % 5 + 1) for i in range(12)])
    kasym = torch.randn(3, 4, generator=g)
    cases = []
    for name, kern, up, down, pad in [
        ("blur_p22", k4, 1, 1, (2, 2)), ("blur_p11", k4, 1, 1, (1, 1)), ("blur_p21", k4, 1, 1, (2, 1)),
        ("blur4_p11", k4 * 4, 1, 1, (1, 1)), ("blur_p00", k4, 1, 1, (0, 0)), ("blur_k3", k3, 1, 1, (1, 1)),
        ("up2_k4", k4 * 4, 2, 1, (2, 1)), ("up2_k2", k2 * 4, 2, 1, (1, 0)), ("down2_k4", k4, 1, 2, (1, 1)),
        ("down2_k2", k2, 1, 2, (0, 0)), ("up2_down2", k4, 2, 2, (2, 2)), ("crop_neg", k4, 1, 1, (-1, 3)),
        ("k12_up2", k12, 2, 1, (6, 5)), ("k12_down2", k12, 1, 2, (5, 5)), ("asym_3x4", kasym, 1, 1, (2, 2)),
        ("up3_down2", k4, 3, 2, (3, 1)),
    ]:
        xin = torch.randn(2, 3, 11, 13, generator=g)
        out = upfirdn2d_native(xin, kern, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        cases.append(dict(name=name, x=xin, kernel=kern.clone(), up=up, down=down, pad=pad, out=out))
    ops["upfirdn2d"] = cases

    mod_cases = []
    for name, kw, shape in [
        ("same", dict(), (2, 6, 9, 9)), ("same_nodemod", dict(demodulate=False), (2, 6, 8, 8)),
        ("up", dict(upsample=True), (2, 6, 7, 7)), ("down", dict(downsample=True), (2, 6, 10, 10)),
        ("k1_nodemod", dict(demodulate=False, _k=1), (2, 6, 5, 5)),
    ]:
        torch.manual_seed(7)
        kk = kw.pop("_k", 3)
        m = RM.ModulatedConv2d(shape[1], 10, kk, 12, **kw)
        m.modulation.weight.data.mul_(0.5)
        xin = torch.randn(*shape).requires_grad_(True)
        st = (torch.rand(shape[0], 12) * 2 - 1).requires_grad_(True)
        out = m(xin, st)
        gy = torch.randn_like(out)
        grads = torch.autograd.grad((out * gy).sum(), [xin, st, m.weight, m.modulation.weight, m.modulation.bias])
        mod_cases.append(dict(name=name, x=xin.detach(), style=st.detach(), gy=gy,
                              sd={k: v.detach().clone() for k, v in m.state_dict().items()},
                              out=out.detach(), grads=[t.detach() for t in grads],
                              demodulate=m.demodulate, upsample=m.upsample, downsample=m.downsample, k=kk))
    ops["modconv"] = mod_cases

    # StyledConv with explicit noise (API-compat extra) and ToRGB
    torch.manual_seed(8)
    scn = RM.StyledConv(6, 10, 3, 12)
    scn.noise.weight.data.fill_(0.3)
    xin, st, nz = torch.randn(2, 6, 8, 8), torch.rand(2, 12) * 2 - 1, torch.randn(2, 1, 8, 8)
    ops["styledconv_noise"] = dict(x=xin, style=st, noise=nz, sd={k: v.clone() for k, v in scn.state_dict().items()},
                                   out=scn(xin, st, noise=nz).detach())
    trgb = RM.ToRGB(6, 12, upsample=True)
    skip = torch.randn(2, 3, 4, 4)
    ops["torgb"] = dict(x=xin, style=st, skip=skip, sd={k: v.clone() for k, v in trgb.state_dict().items()},
                        out=trgb(xin, st, skip).detach())
    torch.save(ops, os.path.join(HERE, "ops.pt"))

    # ------------------------------------------------------------------ 2. small nets
    small = dict(channel=4, structure_channel=8, texture_channel=64, N=2, image_size=256)
    a = ns(**small)
    netg = {}
    torch.manual_seed(11)
    mods = {n: R.init_model(n, a).eval() for n in NET_NAMES if n != "ImageLevelDiscriminator"}
    # Dreal has 26 M parameters whatever `channel` is (models.py:336-346): keep it out of the
    # fixture and pin it through the seeded initialiser instead (seed 12, see contract.json).
    torch.manual_seed(12)
    mods["ImageLevelDiscriminator"] = R.init_model("ImageLevelDiscriminator", a).eval()
    dreal_sha = sha_state(mods["ImageLevelDiscriminator"].state_dict())
    torch.manual_seed(13)
    X = torch.rand(2, 3, 64, 64) * 2 - 1
    Z = torch.rand(2, 2, 4, 4) * 2 - 1
    P = torch.rand(4, 3, 64, 64) * 2 - 1
    Pref = torch.rand(8, 3, 64, 64) * 2 - 1
    X256 = torch.rand(1, 3, 256, 256) * 2 - 1
    with torch.no_grad():
        S, T = mods["DisentanglementEncoder"](X)
        S2 = mods["StructureGenerator"](Z)
        img = mods["Generator"](S2, T)
        zhat = mods["TensorExtractor"](S)
        dreal = mods["ImageLevelDiscriminator"](X256)
        dco, refin = mods["CooccurenceDiscriminator"](P, Pref, ref_batch=2)
        dco2, _ = mods["CooccurenceDiscriminator"](P, ref_input=refin)
        ddist = mods["DistributionDiscriminator"](T)
    netg = dict(cfg=small, dreal_seed=12, dreal_sha=dreal_sha,
                sd={n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in mods.items()
                    if n != "ImageLevelDiscriminator"},
                X=X, Z=Z, P=P, Pref=Pref, X256=X256, S=S, T=T, S2=S2, img=img, zhat=zhat, dreal=dreal,
                dco=dco, refin=refin, dco2=dco2, ddist=ddist)
    torch.save(netg, os.path.join(HERE, "nets_small.pt"))

    # ------------------------------------------------------------------ 3. contract + seeded init
    contract = {}
    for cfgname, kw in [("default256", dict()), ("cfg1_64", dict(image_size=64)), ("small", small),
                        ("N2_512", dict(N=2, image_size=512, channel_multiplier=2))]:
        torch.manual_seed(0)
        entry = {}
        for n in NET_NAMES:
            if n == "CooccurenceDiscriminator" and kw.get("image_size", 256) < 256:
                continue
            sd = R.init_model(n, ns(**kw)).state_dict()
            entry[n] = dict(keys=[[k, list(v.shape)] for k, v in sd.items()], sha256=sha_state(sd))
        contract[cfgname] = dict(cfg=kw, nets=entry)
    json.dump(contract, open(os.path.join(HERE, "contract.json"), "w"), indent=0)

    # ------------------------------------------------------------------ 4. cfg 1
    torch.manual_seed(0)
    a1 = ns(image_size=64)
    E, G = R.init_model("DisentanglementEncoder", a1), R.init_model("Generator", a1)
    Gs, Ex = R.init_model("StructureGenerator", a1), R.init_model("TensorExtractor", a1)
    X = torch.rand(2, 3, 64, 64) * 2 - 1
    with torch.no_grad():
        S1, T1 = E(X)
        Z = torch.rand(2, 1, 4, 4) * 2 - 1
        S2 = Gs(Z)
        Xh = G(S2, T1)
        Sh, _ = E(Xh)
        Zh = Ex(Sh)
    hatM = RU.tensor_to_message(Zh.reshape(2, -1), sigma=1)
    torch.save(dict(X=X, Z=Z, S1=S1, T1=T1, S2=S2, Xh=Xh, Sh=Sh, Zh=Zh, hatM=hatM,
                    sha={"E": sha_state(E.state_dict()), "G": sha_state(G.state_dict()),
                         "Gstru": sha_state(Gs.state_dict()), "Ex": sha_state(Ex.state_dict())}),
               os.path.join(HERE, "cfg1.pt"))

    # ------------------------------------------------------------------ 5. bit path
    bits = {}
    zk = torch.tensor([[-2, -1, -0.75, -0.5, -0.25, -1e-9, -6e-8, -5.96e-8, -1e-7, 0, 1e-9, 0.25, 0.5, 0.75, 1, 1.5]],
                      dtype=torch.float32)
    bits["kat_z"] = zk.tolist()
    bits["kat_decode"] = {str(s): RU.tensor_to_message(zk, s).tolist() for s in (1, 2, 3, 4)}
    Mk = torch.tensor([[0, 0, 0, 1, 1, 0, 1, 1, 0, 1, 1, 0]], dtype=torch.float32)
    bits["kat_m"] = Mk.tolist()
    bits["kat_encode_delta0"] = {str(s): RU.message_to_tensor(Mk, s, 0).tolist() for s in (1, 2, 3)}
    rnd = []
    for s, delta, seed in [(1, 0.5, 0), (2, 0.5, 1), (3, 0.25, 2), (4, 0.5, 3), (1, 0.0, 4)]:
        torch.manual_seed(seed)
        M = torch.randint(0, 2, (4, 48 * s), dtype=torch.float)
        st = torch.get_rng_state()
        Zt = RU.message_to_tensor(M, s, delta)
        torch.set_rng_state(st)
        u = torch.rand(4, 48)                      # the draws rand_like made inside message_to_tensor
        Mh = RU.tensor_to_message(Zt, s)
        rnd.append(dict(sigma=s, delta=delta, M=M.tolist(), u=u.tolist(), Z=Zt.tolist(), Mh=Mh.tolist(),
                        ber=float(torch.mean(torch.abs(M - Mh)))))
    bits["random"] = rnd
    json.dump(bits, open(os.path.join(HERE, "bits.json"), "w"))

    # ------------------------------------------------------------------ 6. two iterations of train.py
    # Shims of SURVEY.md App. D, all outside the reference tree: a stub `dataset` module,
    # Tensor.cuda -> identity (CPU run), float Adam betas (we build the optimisers).
    ds = types.ModuleType("dataset")
    ds.set_dataset = lambda **kw: None
    sys.modules["dataset"] = ds
    import train as RT                      # noqa: E402
    torch.Tensor.cuda = lambda self, *a, **k: self
    ta = ns(channel=4, texture_channel=64, N=1, image_size=256, num_iters=2, start_iter=0, lambda_Ex=10.0,
            lr=0.002, batch_size=2, real_r1=10.0, texture_r1=1.0, dist_r1=1.0, ref_crop=4, n_crop=8,
            d_reg_every=2, log_every=10 ** 9, show_every=10 ** 9, save_every=10 ** 9)
    torch.manual_seed(5)
    random.seed(5)
    order = [("E", "DisentanglementEncoder"), ("G", "Generator"), ("Gstru", "StructureGenerator"),
             ("Ex", "TensorExtractor"), ("Dreal", "ImageLevelDiscriminator"),
             ("Dco", "CooccurenceDiscriminator"), ("Ddist", "DistributionDiscriminator"),
             ("E_ema", "DisentanglementEncoder"), ("G_ema", "Generator"),
             ("Gstru_ema", "StructureGenerator"), ("Ex_ema", "TensorExtractor")]
    trainer = {k: R.init_model(n, ta) for k, n in order}
    for k in ("E", "G", "Gstru", "Ex"):
        trainer[k + "_ema"].eval()
        RU.accumulate(trainer[k + "_ema"], trainer[k], 0)
    init_sha = {k: sha_state(trainer[k].state_dict()) for k, _ in order[:7]}
    P = lambda *ks: [p for k in ks for p in trainer[k].parameters()]  # noqa: E731
    trainer["g_optim"] = torch.optim.Adam(P("E", "G", "Gstru"), lr=ta.lr, betas=(0.0, 0.99))
    trainer["ex_optim"] = torch.optim.Adam(P("Ex"), lr=ta.lr, betas=(0.0, 0.99))
    r = ta.d_reg_every / (ta.d_reg_every + 1)
    trainer["d_optim"] = torch.optim.Adam(P("Dreal", "Dco", "Ddist"), lr=ta.lr * r, betas=(0.0 ** r, 0.99 ** r))
    batches = [torch.rand(2, 3, 256, 256) * 2 - 1 for _ in range(2)]
    RT.train(exp_name="golden", args=ta, loader=batches, trainer=trainer, device="cpu")
    keep = {}
    for k in ("Ex", "Ddist", "Gstru", "Ex_ema"):
        keep[k] = {n: v.clone() for n, v in trainer[k].state_dict().items()}
    for k, names in [("E", ["stem.0.0.weight", "stem.4.conv2.2.bias", "structure.1.0.weight", "texture.3.0.weight"]),
                     ("G", ["layers.0.conv1.conv.weight", "layers.7.conv2.conv.modulation.bias", "to_rgb.0.bias",
                            "layers.4.skip.0.weight"]),
                     ("Dreal", ["convs.0.0.weight", "final_linear.1.weight", "convs.3.conv2.2.bias"]),
                     ("Dco", ["encoder.0.0.weight", "linear.3.weight", "encoder.7.1.bias"]),
                     ("G_ema", ["layers.0.conv1.conv.weight"])]:
        sd = trainer[k].state_dict()
        keep[k] = {n: sd[n].clone() for n in names}
    torch.save(dict(cfg=dict(channel=4, texture_channel=64, N=1, image_size=256), seed=5, d_reg_every=2,
                    note="batches = 2 x (torch.rand(2,3,256,256)*2-1) drawn right after the 11 init_model calls",
                    init_sha=init_sha, after=keep),
               os.path.join(HERE, "train2.pt"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
