"""GPU parity of the op layer: the CUDA kernels, reached through the C ABI
(libideas_b200.so via ctypes), against the CPU oracle (oracle/functional.py) and the
reference-generated goldens (tests/golden/ops.pt).  Tolerance: 1e-3 fp32 relative
(BASELINE.json north_star); the fp32 SIMT path is held to 2e-5, elementwise ops to exact."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import functional as O

import tolerances as T

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north-star tolerance (tcgen05 / TF32 paths)
TOL_FP32 = 3e-5     # fp32 FFMA paths


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


@pytest.fixture(scope="module")
def ops(golden_dir):
    return torch.load(os.path.join(golden_dir, "ops.pt"))


@pytest.fixture(scope="module")
def op():
    from ideas_b200.stylegan2 import op as mod
    return mod


# ------------------------------------------------------------------ A3 fused bias + leaky ReLU
def test_fused_leaky_relu_forward_golden(ops, op):
    for key in ("flrelu", "flrelu_2d"):
        c = ops[key]
        got = op.fused_leaky_relu(c["x"].cuda(), c["bias"].cuda())
        assert torch.equal(got.cpu(), c["out"]), key           # same fp32 op order => bit-exact


@pytest.mark.parametrize("shape", [(2, 5, 7, 9), (3, 8, 16, 16), (4, 64), (2, 128, 33, 31), (1, 3, 5, 5, 2), (0, 8, 4, 4)])
def test_fused_leaky_relu_first_and_second_order(op, shape):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    gy = torch.randn(*shape, generator=g)

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        bb = b.to(dev).requires_grad_(True)
        out = fn(xx, bb)
        gx, gb = torch.autograd.grad(out, [xx, bb], gy.to(dev))
        return out, gx, gb

    o_c, gx_c, gb_c = run(lambda a, c: O.fused_leaky_relu(a, c), "cpu")
    o_g, gx_g, gb_g = run(lambda a, c: op.fused_leaky_relu(a, c), "cuda")
    assert torch.equal(o_g.cpu(), o_c.detach())
    if x.numel():
        assert rel(gx_g, gx_c) <= 1e-6
        assert rel(gb_g, gb_c) <= 1e-5


def test_fused_leaky_relu_double_backward_matches_oracle(op):
    # R1: grad of (dL/dx) w.r.t. the upstream gradient -- fused_act.py:43-49
    g = torch.Generator().manual_seed(2)
    x, b = torch.randn(2, 8, 6, 6, generator=g), torch.randn(8, generator=g)

    def run(fn, dev):
        xx = x.to(dev)
        gy = torch.randn(2, 8, 6, 6, generator=torch.Generator().manual_seed(3)).to(dev).requires_grad_(True)
        xin = xx.clone().requires_grad_(True)
        out = fn(xin, b.to(dev))
        (gx,) = torch.autograd.grad(out, xin, gy, create_graph=True)
        (ggy,) = torch.autograd.grad(gx.pow(2).sum(), gy)
        return ggy

    assert rel(run(op.fused_leaky_relu, "cuda"), run(O.fused_leaky_relu, "cpu")) <= 1e-6


def test_native_bias_act_codes(op):
    """The C entry point honours every act*10+grad code of fused_bias_act_kernel.cu:36-47."""
    from ideas_b200.stylegan2.op.fused_act import _bias_act
    g = torch.Generator().manual_seed(4)
    x, b, r = torch.randn(2, 8, 4, 4, generator=g), torch.randn(8, generator=g), torch.randn(2, 8, 4, 4, generator=g)
    for act, grad in [(1, 0), (1, 1), (1, 2), (3, 0), (3, 1), (3, 2)]:
        want = O.bias_act(x, b, r, act, grad, 0.2, 1.5)
        xc = x.cuda().contiguous(memory_format=torch.channels_last)
        rc = r.cuda().contiguous(memory_format=torch.channels_last)
        got = _bias_act(xc, b.cuda(), rc, act, grad, 0.2, 1.5)
        assert torch.equal(got.cpu(), want), (act, grad)


def test_ops_refuse_cpu_tensors(op):
    with pytest.raises(RuntimeError):
        op.fused_leaky_relu(torch.randn(1, 4, 2, 2), torch.zeros(4))
    with pytest.raises(RuntimeError):
        op.upfirdn2d(torch.randn(1, 4, 8, 8), O.make_kernel([1, 3, 3, 1]), pad=(2, 1))


# ------------------------------------------------------------------ A4 upfirdn2d
def test_upfirdn2d_golden_all_modes(ops, op):
    for c in ops["upfirdn2d"]:
        got = op.upfirdn2d(c["x"].cuda(), c["kernel"].cuda(), up=c["up"], down=c["down"], pad=tuple(c["pad"]))
        assert rel(got, c["out"]) <= 2e-6, c["name"]


@pytest.mark.parametrize("C,H,W,pad,gain", [(8, 16, 16, (2, 2), 1), (128, 33, 33, (1, 1), 4), (4, 9, 31, (2, 1), 1),
                                            (12, 64, 64, (1, 1), 1), (3, 17, 13, (2, 2), 1), (64, 257, 257, (1, 1), 4)])
def test_blur_fwd_bwd_double_bwd(op, C, H, W, pad, gain):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, C, H, W, generator=g)
    k = O.make_kernel([1, 3, 3, 1]) * gain

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        out = fn(xx, k.to(dev), pad=pad)
        gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(6)).to(dev).requires_grad_(True)
        (gx,) = torch.autograd.grad(out, xx, gy, create_graph=True)
        (ggy,) = torch.autograd.grad(gx.pow(2).sum(), gy)
        return out, gx, ggy

    for a, b in zip(run(op.upfirdn2d, "cuda"), run(O.upfirdn2d, "cpu")):
        assert rel(a, b) <= 5e-6


@pytest.mark.parametrize("C,H,W,up,down,pad", [
    (8, 16, 16, 1, 2, (1, 1)), (64, 255, 255, 1, 2, (1, 1)), (12, 33, 20, 1, 2, (2, 1)), (4, 9, 31, 1, 2, (0, 0)),
    (128, 64, 64, 1, 2, (2, 2)), (8, 16, 16, 2, 1, (2, 1)), (64, 64, 64, 2, 1, (2, 1)), (12, 17, 9, 2, 1, (1, 1)),
    (4, 5, 31, 2, 1, (3, 2)), (32, 31, 33, 2, 1, (2, 2)), (3, 12, 12, 2, 1, (2, 1)), (6, 12, 12, 1, 2, (1, 1))])
def test_resampling_fast_paths_fwd_bwd_double_bwd(op, C, H, W, up, down, pad):
    """upfirdn2d with up = 2 or down = 2 (the fused skip-branch kernels; C % 4 != 0 exercises the generic kernel)
    against the oracle restatement of upfirdn2d_native (upfirdn2d.py:159-200), forward, backward (which runs the
    opposite resampling kernel) and double backward."""
    g = torch.Generator().manual_seed(15)
    x = torch.randn(2, C, H, W, generator=g)
    k = O.make_kernel([1, 3, 3, 1]) * (up ** 2)

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        out = fn(xx, k.to(dev), up=up, down=down, pad=pad)
        gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(16)).to(dev).requires_grad_(True)
        (gx,) = torch.autograd.grad(out, xx, gy, create_graph=True)
        (ggy,) = torch.autograd.grad(gx.pow(2).sum(), gy)
        return out, gx, ggy

    for a, b in zip(run(op.upfirdn2d, "cuda"), run(O.upfirdn2d, "cpu")):
        assert a.shape == b.shape
        assert rel(a, b) <= 5e-6


def test_resampling_non_separable_taps(op):
    g = torch.Generator().manual_seed(17)
    x = torch.randn(2, 8, 13, 11, generator=g)
    k = torch.randn(4, 4, generator=g)
    for up, down, pad in ((1, 2, (1, 1)), (2, 1, (2, 1))):
        got = op.upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad)
        assert rel(got, O.upfirdn2d(x, k, up=up, down=down, pad=pad)) <= 5e-6


def test_blur_bias_act_fused_epilogue(op):
    g = torch.Generator().manual_seed(7)
    x, b = torch.randn(2, 16, 12, 12, generator=g), torch.randn(16, generator=g)
    k = O.make_kernel([1, 3, 3, 1]) * 4

    def run(dev, fused):
        xx, bb = x.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
        if fused:
            out = op.blur_bias_act(xx, k.to(dev), (1, 1), bb)
        else:
            out = O.fused_leaky_relu(O.upfirdn2d(xx, k, pad=(1, 1)), bb)
        gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(8)).to(dev)
        gx, gb = torch.autograd.grad(out, [xx, bb], gy)
        return out, gx, gb

    for a, c in zip(run("cuda", True), run("cpu", False)):
        assert rel(a, c) <= 1e-5


def test_upfirdn2d_empty_and_identity(op):
    k1 = torch.ones(1, 1).cuda()
    x = torch.randn(1, 2, 5, 5).cuda()
    assert torch.equal(op.upfirdn2d(x, k1), x)
    assert op.upfirdn2d(torch.zeros(0, 4, 5, 5).cuda(), O.make_kernel([1, 3, 3, 1]).cuda(), pad=(2, 1)).shape == (0, 4, 5, 5)


# ------------------------------------------------------------------ A5 convolution family
CONV_CASES = [
    # N, C, K, H, W, k, stride, pad
    (2, 3, 32, 16, 16, 1, 1, 0), (2, 32, 64, 18, 18, 3, 1, 0), (2, 64, 64, 17, 17, 3, 2, 0), (1, 8, 128, 16, 16, 3, 1, 1),
    (2, 33, 7, 9, 11, 3, 1, 1), (2, 64, 32, 15, 15, 1, 2, 0), (3, 96, 48, 2, 2, 2, 1, 0), (2, 128, 3, 8, 8, 1, 1, 0),
    (1, 128, 128, 33, 33, 3, 2, 0), (2, 256, 512, 8, 8, 3, 1, 1),
    # 1x1 convs with <= 4 channels on one side: the HBM-bound pointwise kernels (image ends of the nets)
    (2, 3, 64, 64, 64, 1, 1, 0), (2, 128, 3, 32, 32, 1, 1, 0), (3, 1, 32, 16, 16, 1, 1, 0), (2, 32, 1, 16, 16, 1, 1, 0),
    (2, 4, 256, 9, 7, 1, 1, 0), (1, 2048, 2, 5, 5, 1, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_dgrad_wgrad(case, conv_mode):
    from ideas_b200.stylegan2.op import conv as C
    N, Ci, K, H, W, k, stride, pad = case
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, Ci, H, W, generator=g)
    w = torch.randn(K, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5
    b = torch.randn(K, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = O.fused_leaky_relu(F.conv2d(xr, wr, None, stride, pad), br)
    gy = torch.randn(want.shape, generator=g)
    wg = torch.autograd.grad(want, [xr, wr, br], gy)
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    wp = C.PackWeight.apply(wc, False, 1.0)
    got = C.conv2d(xc, wp, bc, K=K, kh=k, kw=k, stride=stride, pad=pad, act=True)
    gg = torch.autograd.grad(got, [xc, wc, bc], gy.cuda())
    assert T.rel(got, want) <= T.OUT[conv_mode]
    for name, a, c in zip(("dx", "dw", "db"), gg, wg):
        T.assert_grad_through_act(a, c, conv_mode, name)


@pytest.mark.parametrize("case", [(2, 16, 8, 7, 7, 1, 2), (2, 64, 32, 16, 16, 3, 2), (1, 32, 32, 5, 6, 3, 1), (2, 128, 64, 16, 16, 1, 2)])
def test_conv_transpose2d(case, conv_mode):
    from ideas_b200.stylegan2.op import conv as C
    N, Ci, Co, H, W, k, stride = case
    g = torch.Generator().manual_seed(10)
    x = torch.randn(N, Ci, H, W, generator=g)
    w = torch.randn(Ci, Co, k, k, generator=g) / (Ci * k * k) ** 0.5
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    want = F.conv_transpose2d(xr, wr, stride=stride)
    gy = torch.randn(want.shape, generator=g)
    wg = torch.autograd.grad(want, [xr, wr], gy)
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    wp = C.PackWeight.apply(wc, False, 1.0)
    got = C.conv_transpose2d(xc, wp, C_out=Co, kh=k, kw=k, stride=stride, pad=0)
    gg = torch.autograd.grad(got, [xc, wc], gy.cuda())
    assert T.rel(got, want) <= T.OUT[conv_mode]
    for a, c in zip(gg, wg):
        assert T.rel(a, c) <= T.GRAD[conv_mode]


def test_conv_second_order_r1_style(conv_mode):
    """R1 differentiates dD/dx w.r.t. the weights: conv double-backward (utils.py:112-118)."""
    from ideas_b200.stylegan2.op import conv as C
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 8, 10, 10, generator=g)
    w1 = torch.randn(16, 8, 3, 3, generator=g) / 8.5
    w2 = torch.randn(4, 16, 3, 3, generator=g) / 12
    b1 = torch.randn(16, generator=g)

    def oracle():
        xr = x.clone().requires_grad_(True)
        a, c = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
        bb = b1.clone().requires_grad_(True)
        y = F.conv2d(O.fused_leaky_relu(F.conv2d(xr, a, None, 1, 1), bb), c, None, 2, 0)
        (gx,) = torch.autograd.grad(y.sum(), xr, create_graph=True)
        pen = gx.pow(2).sum()
        return (pen,) + torch.autograd.grad(pen, [a, c])

    def ours():
        xr = x.cuda().requires_grad_(True)
        a, c = w1.cuda().requires_grad_(True), w2.cuda().requires_grad_(True)
        bb = b1.cuda().requires_grad_(True)
        h = C.conv2d(xr, C.PackWeight.apply(a, False, 1.0), bb, K=16, kh=3, kw=3, stride=1, pad=1, act=True)
        y = C.conv2d(h, C.PackWeight.apply(c, False, 1.0), None, K=4, kh=3, kw=3, stride=2, pad=0)
        (gx,) = torch.autograd.grad(y.sum(), xr, create_graph=True)
        pen = gx.pow(2).sum()
        return (pen,) + torch.autograd.grad(pen, [a, c])

    for name, a, c in zip(("penalty", "dw1", "dw2"), ours(), oracle()):
        T.assert_grad_through_act(a, c, conv_mode, name)


# ------------------------------------------------------------------ A1/A2 modulated convolution
def _load_modconv(c):
    from ideas_b200.stylegan2 import model as M
    sd = c["sd"]
    cout, cin = sd["weight"].shape[1], sd["weight"].shape[2]
    m = M.ModulatedConv2d(cin, cout, c["k"], sd["modulation.weight"].shape[1], demodulate=c["demodulate"],
                          upsample=c["upsample"], downsample=c["downsample"])
    m.load_state_dict(sd)
    return m.cuda()


def test_modulated_conv_golden(ops, conv_mode):
    for c in ops["modconv"]:
        m = _load_modconv(c)
        x = c["x"].cuda().requires_grad_(True)
        st = c["style"].cuda().requires_grad_(True)
        out = m(x, st)
        assert T.rel(out, c["out"]) <= T.OUT[conv_mode], c["name"]
        grads = torch.autograd.grad((out * c["gy"].cuda()).sum(), [x, st, m.weight, m.modulation.weight, m.modulation.bias])
        for i, (a, b) in enumerate(zip(grads, c["grads"])):
            assert T.rel(a, b) <= T.GRAD[conv_mode], (c["name"], i)


@pytest.mark.parametrize("cin,cout,res,up", [(64, 64, 16, False), (64, 32, 8, True), (128, 128, 12, False), (32, 64, 9, True)])
def test_styled_conv_vs_oracle(cin, cout, res, up, conv_mode):
    """StyledConv_without_noise (the IDEAS block) at tensor-core-eligible widths."""
    from ideas_b200.stylegan2 import model as M
    torch.manual_seed(12)
    m = M.StyledConv_without_noise(cin, cout, 3, 48, upsample=up)
    m.activate.bias.data.normal_()
    x = torch.randn(2, cin, res, res)
    st = torch.rand(2, 48) * 2 - 1
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "kernel" not in k) for k, v in m.state_dict().items()}
    xr, sr = x.clone().requires_grad_(True), st.clone().requires_grad_(True)
    want = O.styled_conv(xr, sr, sd["conv.weight"], sd["conv.modulation.weight"], sd["conv.modulation.bias"],
                         sd["activate.bias"], upsample=up, blur_kernel=sd.get("conv.blur.kernel"))
    gy = torch.randn_like(want)
    names = ["conv.weight", "conv.modulation.weight", "conv.modulation.bias", "activate.bias"]
    wg = torch.autograd.grad(want, [xr, sr] + [sd[n] for n in names], gy)
    m = m.cuda()
    xc, sc = x.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
    got = m(xc, sc)
    params = dict(m.named_parameters())
    gg = torch.autograd.grad(got, [xc, sc] + [params[n] for n in names], gy.cuda())
    assert T.rel(got, want) <= T.OUT[conv_mode]
    for name, a, b in zip(["dx", "dstyle"] + names, gg, wg):
        T.assert_grad_through_act(a, b, conv_mode, name)


def test_styled_conv_noise_and_torgb_golden(ops, conv_mode):
    from ideas_b200.stylegan2 import model as M
    c = ops["styledconv_noise"]
    m = M.StyledConv(6, 10, 3, 12)
    m.load_state_dict(c["sd"])
    m = m.cuda()
    assert rel(m(c["x"].cuda(), c["style"].cuda(), noise=c["noise"].cuda()), c["out"]) <= TOL
    c = ops["torgb"]
    t = M.ToRGB(6, 12, upsample=True)
    t.load_state_dict(c["sd"])
    t = t.cuda()
    assert rel(t(c["x"].cuda(), c["style"].cuda(), c["skip"].cuda()), c["out"]) <= TOL


# ------------------------------------------------------------------ A10 patchify (utils.py:127-149)
@pytest.mark.parametrize("B,C,H,W,n_crop", [(2, 3, 256, 256, 8), (3, 3, 64, 96, 5), (1, 4, 128, 128, 32)])
def test_patchify_matches_interpolate_loop(B, C, H, W, n_crop):
    import random
    from ideas_b200 import utils as U
    from oracle.train_step import draw_crops, patchify_image as oracle_patchify
    torch.manual_seed(3)
    random.seed(3)
    crops = draw_crops(n_crop, H, W)
    crops[0] = (0, 0, H // 4, W // 4)                       # identity-size crop: exact copy
    crops[-1] = (H - H // 8, W - W // 8, H // 8, W // 8)    # smallest crop in the far corner (2x upsampling)
    img = torch.rand(B, C, H, W) * 2 - 1
    ir = img.clone().requires_grad_(True)
    want = oracle_patchify(ir, crops)
    gy = torch.randn_like(want)
    (wg,) = torch.autograd.grad(want, ir, gy)
    ic = img.cuda().requires_grad_(True)
    got = U.patchify_image(ic, n_crop, crops=crops)
    (gg,) = torch.autograd.grad(got, ic, gy.cuda())
    assert got.shape == want.shape
    assert rel(got, want) <= 2e-6
    assert rel(gg, wg) <= 1e-5
    # device-resident boxes give the same result (CUDA-graph path)
    got2 = U.patchify_image(ic, n_crop, crops=U.crops_to_device(crops, "cuda"))
    assert torch.equal(got2, got)


@pytest.mark.parametrize("C,K,H", [(32, 64, 20), (64, 128, 33), (32, 384, 12), (64, 1024, 9), (8, 20, 16)])
def test_conv_act_blur_fused_backward(C, K, H):
    """ConvActBlur (conv -> FusedLeakyReLU -> Blur as one node, backward = one blur^T * mask kernel) against the
    same three ops as separate autograd nodes; K = 384 / 20 take the documented fallback inside the node.  Also the
    create_graph path (what the R1 penalty uses) against the unfused double backward."""
    from ideas_b200 import _lib as L
    from ideas_b200.stylegan2.op import conv as CV
    from ideas_b200.stylegan2.op import fused_leaky_relu, upfirdn2d
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(2, C, H, H, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    wp = torch.randn(9, K, C, device="cuda", generator=g) / (9 * C) ** 0.5
    b = torch.randn(K, device="cuda", generator=g)
    k = (O.make_kernel([1, 3, 3, 1])).cuda()
    gz = torch.randn(2, K, H + 1, H + 1, device="cuda", generator=g)
    geom = CV.Geom.forward(x.shape, K, 3, 3, 1, 1)

    def fused(xx, ww, bb):
        return CV.ConvActBlur.apply(xx, ww, bb, geom, 0.2, 2 ** 0.5, k, (2, 2))

    def split(xx, ww, bb):
        y = CV.ConvFwd.apply(xx, ww, bb, geom, L.ACT_LRELU, 0.2, 2 ** 0.5)
        return upfirdn2d(y, k, pad=(2, 2))

    outs = []
    for fn in (fused, split):
        xx, ww, bb = x.clone().requires_grad_(True), wp.clone().requires_grad_(True), b.clone().requires_grad_(True)
        z = fn(xx, ww, bb)
        gx, gw, gb = torch.autograd.grad(z, [xx, ww, bb], gz)
        # second order: d/dw of |dz/dx . gz|^2, as in an R1 penalty
        xx2, ww2 = x.clone().requires_grad_(True), wp.clone().requires_grad_(True)
        z2 = fn(xx2, ww2, b)
        (g1,) = torch.autograd.grad(z2.sum(), xx2, create_graph=True)
        (ggw,) = torch.autograd.grad(g1.pow(2).sum(), ww2)
        outs.append((z, gx, gw, gb, ggw))
    for a, c, name in zip(outs[0], outs[1], ("z", "gx", "gw", "gb", "ggw")):
        assert rel(a, c) <= 2e-5, (name, rel(a, c))


@pytest.mark.parametrize("C,H,W,pad", [(32, 16, 16, 1), (64, 9, 31, 1), (8, 5, 4, 2), (3, 7, 6, 1), (4, 2, 3, 1), (128, 64, 64, 1)])
def test_reflect_pad_nhwc_fwd_bwd_double_bwd(C, H, W, pad):
    """NHWC reflection padding against torch's F.pad(mode="reflect") on the CPU (nn.ReflectionPad2d of the
    reference, models.py:102-108): forward, gradient (adjoint gather) and double backward; bit-exact (pure copies
    and <= 9-term sums in a fixed order can differ only by summation order)."""
    from ideas_b200.stylegan2.op.elementwise import reflect_pad
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, C, H, W, generator=g)

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        out = fn(xx)
        gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(42)).to(dev).requires_grad_(True)
        (gx,) = torch.autograd.grad(out, xx, gy, create_graph=True)
        (ggy,) = torch.autograd.grad(gx.pow(2).sum(), gy)
        return out, gx, ggy

    got = run(lambda t: reflect_pad(t, pad), "cuda")
    want = run(lambda t: torch.nn.functional.pad(t, (pad,) * 4, mode="reflect"), "cpu")
    assert torch.equal(got[0].cpu(), want[0])
    for a, b in zip(got[1:], want[1:]):
        assert a.shape == b.shape and rel(a, b) <= 1e-6
    assert got[0].is_contiguous(memory_format=torch.channels_last) or C == 1


@pytest.mark.parametrize("cin,cout,res,up", [(32, 64, 8, False), (64, 32, 8, True)])
def test_styled_conv_path_length_second_order(cin, cout, res, up):
    """Path-length regularisation (stylegan2/train.py:85-98): the gradient of <y, noise> w.r.t.
    the style, taken with create_graph=True, squared and back-propagated into the layer's parameters and input --
    i.e. the double backward of the modulated convolution -- against the oracle on the CPU.  Runs on the exact-fp32
    FFMA kernels so the comparison is tight."""
    from ideas_b200 import _lib as L
    from ideas_b200.stylegan2 import model as M
    from ideas_b200.stylegan2.op import conv as CV
    old = CV.set_default_impl(L.IMPL_SIMT)
    try:
        torch.manual_seed(21)
        m = M.StyledConv_without_noise(cin, cout, 3, 24, upsample=up)
        m.activate.bias.data.normal_()
        x = torch.randn(2, cin, res, res)
        st = torch.rand(2, 24) * 2 - 1
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "kernel" not in k)
              for k, v in m.state_dict().items()}
        names = ["conv.weight", "conv.modulation.weight", "conv.modulation.bias"]

        def pl(fn_out, style, leaves):
            noise = torch.randn(fn_out.shape, generator=torch.Generator().manual_seed(22)).to(fn_out.device)
            (g,) = torch.autograd.grad((fn_out * noise).sum(), style, create_graph=True)
            penalty = g.pow(2).sum(1).mean()
            return penalty, torch.autograd.grad(penalty, leaves)

        xr, sr = x.clone().requires_grad_(True), st.clone().requires_grad_(True)
        want = O.styled_conv(xr, sr, sd["conv.weight"], sd["conv.modulation.weight"], sd["conv.modulation.bias"],
                             sd["activate.bias"], upsample=up, blur_kernel=sd.get("conv.blur.kernel"))
        p_want, g_want = pl(want, sr, [xr] + [sd[n] for n in names])
        m = m.cuda()
        xc, sc = x.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
        params = dict(m.named_parameters())
        p_got, g_got = pl(m(xc, sc), sc, [xc] + [params[n] for n in names])
        assert abs(float(p_got) - float(p_want)) <= 1e-4 * abs(float(p_want))
        for name, a, b in zip(["dx"] + names, g_got, g_want):
            assert rel(a, b) <= 2e-3, (name, rel(a, b))      # gradients through the leaky-ReLU mask: a few flips
    finally:
        CV.set_default_impl(old)


@pytest.mark.parametrize("cin,cout,res,up", [(64, 64, 16, False), (64, 32, 16, True)])
def test_styled_conv_path_length_second_order_tcgen05(cin, cout, res, up):
    """The same double backward with every convolution of the first AND second order graph on the tcgen05 kernels
    (tf32 operands): the penalty and its parameter / input gradients against the oracle, at the tf32 bounds of
    tests/tolerances.py (gradients cross a leaky-ReLU mask twice)."""
    from ideas_b200 import _lib as L
    from ideas_b200.stylegan2 import model as M
    from ideas_b200.stylegan2.op import conv as CV
    from tolerances import record
    old = CV.set_default_impl(L.IMPL_AUTO)
    try:
        torch.manual_seed(23)
        m = M.StyledConv_without_noise(cin, cout, 3, 24, upsample=up)
        m.activate.bias.data.normal_()
        x = torch.randn(2, cin, res, res)
        st = torch.rand(2, 24) * 2 - 1
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "kernel" not in k)
              for k, v in m.state_dict().items()}
        names = ["conv.weight", "conv.modulation.weight", "conv.modulation.bias"]

        def pl(fn_out, style, leaves):
            noise = torch.randn(fn_out.shape, generator=torch.Generator().manual_seed(24)).to(fn_out.device)
            (g,) = torch.autograd.grad((fn_out * noise).sum(), style, create_graph=True)
            penalty = g.pow(2).sum(1).mean()
            return penalty, torch.autograd.grad(penalty, leaves)

        xr, sr = x.clone().requires_grad_(True), st.clone().requires_grad_(True)
        want = O.styled_conv(xr, sr, sd["conv.weight"], sd["conv.modulation.weight"], sd["conv.modulation.bias"],
                             sd["activate.bias"], upsample=up, blur_kernel=sd.get("conv.blur.kernel"))
        p_want, g_want = pl(want, sr, [xr] + [sd[n] for n in names])
        m = m.cuda()
        xc, sc = x.cuda().requires_grad_(True), st.cuda().requires_grad_(True)
        params = dict(m.named_parameters())
        n0 = L.launch_count()
        p_got, g_got = pl(m(xc, sc), sc, [xc] + [params[n] for n in names])
        assert L.launch_count() > n0
        errs = {"penalty": abs(float(p_got) - float(p_want)) / abs(float(p_want))}
        for name, a, b in zip(["dx"] + names, g_got, g_want):
            a, b = a.detach().cpu().double(), b.detach().double()
            errs[name] = float((a - b).norm() / b.norm())
        for name, e in errs.items():
            record(f"pl_reg_tcgen05_{'up_' if up else ''}{name}", e)
        # measured on B200: penalty 1.2e-3 (5.8e-3 up-sampling), gradients <= 1.6e-2 relative L2; bounds ~2x that
        assert errs.pop("penalty") <= 1.2e-2, errs
        assert max(errs.values()) <= 3e-2, errs
    finally:
        CV.set_default_impl(old)


# ----------------------------------------------------------------------------------- A6 EqualLinear GEMM
@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 7), (32, 512, 2048), (96, 1, 512), (128, 512, 8192), (33, 130, 70)])
def test_matmul_nt_matches_fp64_any_strides(shape):
    """ideas_gemm_nt through op/linear.py: exact-fp32 a @ b^T for contiguous, transposed and broadcast operands,
    including the split reduction (8192 -> 512 head)."""
    from ideas_b200.stylegan2.op.linear import matmul_nt
    M, N, R = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, R, device="cuda", generator=g)
    b = torch.randn(N, R, device="cuda", generator=g)
    want = (a.double() @ b.double().t())
    scale = want.abs().max().clamp_min(1e-9)
    assert float((matmul_nt(a, b).double() - want).abs().max() / scale) <= 2e-6
    at, bt = a.t().contiguous().t(), b.t().contiguous().t()            # same values, reduction index strided
    assert float((matmul_nt(at, bt).double() - want).abs().max() / scale) <= 2e-6
    ones = torch.ones(1, 1, device="cuda").expand(M, R)                # stride-0 operand (gradient of a sum)
    assert float((matmul_nt(ones, b).double() - ones.double() @ b.double().t()).abs().max()) <= 2e-6 * R ** 0.5 * 4


def test_equal_linear_first_and_second_order_vs_oracle():
    """EqualLinear (fused_lrelu head) on the hand-written GEMM: value, gradients and the R1-style double backward
    (gradient of |d out / d x|^2 w.r.t. the weights) against the oracle's torch formulation."""
    from ideas_b200.stylegan2.model import EqualLinear
    torch.manual_seed(3)
    lin = EqualLinear(96, 40, activation="fused_lrelu", lr_mul=0.5)
    lin.bias.data.normal_()
    x = torch.randn(7, 96)
    w, b = lin.weight.detach().clone().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    yr = O.equal_linear(xr, w, b, lr_mul=0.5, activation="fused_lrelu")
    (gr,) = torch.autograd.grad(yr.sum(), xr, create_graph=True)
    pen_r = gr.pow(2).sum()
    want = torch.autograd.grad(pen_r + yr.pow(2).sum(), [w, b, xr])
    lin = lin.cuda()
    xc = x.cuda().requires_grad_(True)
    yc = lin(xc)
    (gc,) = torch.autograd.grad(yc.sum(), xc, create_graph=True)
    got = torch.autograd.grad(gc.pow(2).sum() + yc.pow(2).sum(), [lin.weight, lin.bias, xc])
    assert rel(yc, yr) <= 2e-6
    for a, c, name in zip(got, want, ("dW", "db", "dx")):
        assert rel(a, c) <= 2e-5, (name, rel(a, c))


@pytest.mark.parametrize("shape,pad", [((2, 8, 9, 7), (2, 0)), ((3, 64, 16, 16), (2, 0)), ((2, 12, 5, 6), (1, 1))])
def test_upfirdn2d_up2_residual_merge(op, shape, pad):
    """(upfirdn2d(x, up=2) + residual) * s in the up-sampling kernel's epilogue (ideas_upfirdn2d_res) vs the two-step
    composition, values and both gradients."""
    from ideas_b200.stylegan2.model import make_kernel
    g = torch.Generator().manual_seed(2)
    k = (make_kernel([1, 3, 3, 1]) * 4).cuda()
    x = torch.randn(*shape, generator=g).cuda().requires_grad_(True)
    want0 = op.upfirdn2d(x, k, up=2, pad=pad)
    res = torch.randn(*want0.shape, generator=g).cuda().requires_grad_(True)
    s = 2 ** -0.5
    want = (want0 + res) * s
    got = op.upfirdn2d(x, k, up=2, pad=pad, residual=res, res_scale=s)
    gy = torch.randn(*want.shape, generator=g).cuda()
    wg = torch.autograd.grad(want, [x, res], gy)
    gg = torch.autograd.grad(got, [x, res], gy)
    assert rel(got, want) <= 1e-6
    for a, b in zip(gg, wg):
        assert rel(a, b) <= 1e-6


def test_conv2d_residual_merge_autograd(conv_mode):
    """conv2d(..., residual=r, res_scale=s) == (conv2d(..) + r) * s with gradients to x, w, bias and r, first and second
    order (the discriminators' ResBlocks are differentiated twice by R1)."""
    from ideas_b200.stylegan2.op import conv as C
    g = torch.Generator().manual_seed(4)
    N, Cc, K, H = 2, 32, 64, 12
    x = torch.randn(N, Cc, H, H, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(K, Cc, 1, 1, generator=g) / Cc ** 0.5).cuda().requires_grad_(True)
    b = torch.randn(K, generator=g).cuda().requires_grad_(True)
    r = torch.randn(N, K, H, H, generator=g).cuda().requires_grad_(True)
    s = 2 ** -0.5

    def run(fused):
        wp = C.PackWeight.apply(w, False, 1.0)
        if fused:
            y = C.conv2d(x, wp, b, K=K, kh=1, kw=1, residual=r, res_scale=s)
        else:
            y = (C.conv2d(x, wp, b, K=K, kh=1, kw=1) + r) * s
        (gx,) = torch.autograd.grad(y.pow(2).sum(), x, create_graph=True)
        pen = gx.pow(2).sum()
        return y, torch.autograd.grad(pen + y.sum(), [x, w, b, r])

    y0, g0 = run(False)
    y1, g1 = run(True)
    tol = 1e-5 if conv_mode == "fp32" else 2e-3
    assert T.rel(y1, y0) <= tol
    for a, c in zip(g1, g0):
        assert T.rel(a, c) <= tol * 5


@pytest.mark.parametrize("up", [False, True])
def test_styled_conv_pair_with_premodulation(conv_mode, up):
    """conv1 -> conv2 of a StyledResBlock with conv2's style folded into conv1's output (post / premodulated path of
    op/conv.py) against the plain composition of the two modules: output and every gradient, incl. both styles."""
    from ideas_b200.stylegan2 import model as M
    torch.manual_seed(11)
    c1 = M.StyledConv_without_noise(64, 64, 3, 24, upsample=up).cuda()
    c2 = M.StyledConv_without_noise(64, 96, 3, 24).cuda()
    for c in (c1, c2):
        c.activate.bias.data.normal_()
    x = torch.randn(2, 64, 12, 12).cuda().requires_grad_(True)
    st = (torch.rand(2, 24) * 2 - 1).cuda().requires_grad_(True)
    params = [x, st] + list(c1.parameters()) + list(c2.parameters())
    plain = c2(c1(x, st), st)
    gy = torch.randn_like(plain)
    want = torch.autograd.grad(plain, params, gy)
    s2 = c2.conv.modulation(st)
    a, am = c1(x, st, post_modulation=s2)
    fused = c2(am, st, modulation=s2, premodulated=True)
    got = torch.autograd.grad(fused, params, gy)
    tol = 2e-5 if conv_mode == "fp32" else 1e-3
    assert T.rel(fused, plain) <= tol
    assert T.rel(a, c1(x, st)) <= tol
    for g_, w_, p_ in zip(got, want, params):
        # same kernels and same mask (both sides run the same forward), so no flip noise: plain relative error
        assert T.rel(g_, w_) <= 5 * tol, tuple(p_.shape)
    # both outputs of conv1 used: the general path
    s2 = c2.conv.modulation(st)
    a, am = c1(x, st, post_modulation=s2)
    both = torch.autograd.grad((c2(am, st, modulation=s2, premodulated=True) * gy).sum() + (a * a).sum(), params)
    ref = torch.autograd.grad((c2(c1(x, st), st) * gy).sum() + (c1(x, st) ** 2).sum(), params)
    for g_, w_ in zip(both, ref):
        assert T.rel(g_, w_) <= 5 * tol


def test_split_down_matches_separate_ops():
    """SplitDown (op/upfirdn2d.py): (x, blur_down2(x)) as one node whose backward adds the two gradients inside the
    blur's adjoint kernel, against the two separate uses of x that autograd would sum."""
    from ideas_b200.stylegan2.model import make_kernel
    from ideas_b200.stylegan2.op.upfirdn2d import SplitDown, upfirdn2d
    g = torch.Generator().manual_seed(6)
    k = make_kernel([1, 3, 3, 1]).cuda()
    for shape, pad in (((2, 8, 16, 16), (1, 1)), ((3, 64, 33, 21), (1, 1)), ((2, 12, 9, 10), (2, 1))):
        x = torch.randn(*shape, generator=g).cuda().requires_grad_(True)
        xa, low = SplitDown.apply(x, k, (pad[0], pad[1], pad[0], pad[1]))
        want_low = upfirdn2d(x, k, down=2, pad=pad)
        assert rel(low, want_low) <= 1e-6 and torch.equal(xa, x)
        g_main = torch.randn(*shape, generator=g).cuda()
        g_low = torch.randn(*low.shape, generator=g).cuda()
        (got,) = torch.autograd.grad([xa, low], [x], [g_main, g_low])
        (want,) = torch.autograd.grad([x * 1.0, want_low], [x], [g_main, g_low])
        assert rel(got, want) <= 1e-6, shape
        # only one of the two outputs used
        xa, low = SplitDown.apply(x, k, (pad[0], pad[1], pad[0], pad[1]))
        (g1,) = torch.autograd.grad(low, [x], g_low)
        (w1,) = torch.autograd.grad(upfirdn2d(x, k, down=2, pad=pad), [x], g_low)
        assert rel(g1, w1) <= 1e-6


@pytest.mark.parametrize("down", [False, True])
def test_resblock_scale_folding_matches_unfused(conv_mode, down):
    """ResBlock with (out + skip)/sqrt(2) folded into conv2's activation gain and the skip convolution's weights (and,
    for down-sampling blocks, SplitDown at the input) against the same block evaluated layer by layer with an explicit
    add_scale merge: output and every gradient."""
    import math
    from ideas_b200.models import ResBlock
    from ideas_b200.stylegan2.op.elementwise import add_scale
    torch.manual_seed(5)
    blk = ResBlock(32, 64, downsample=down).cuda()
    for p in blk.parameters():
        if p.dim() == 1:
            p.data.normal_()
    x = torch.randn(2, 32, 20, 20).cuda().requires_grad_(True)
    params = [x] + list(blk.parameters())
    got = blk(x)
    plain = add_scale(blk.conv2(blk.conv1(x)), blk.skip(x), 1 / math.sqrt(2))
    gy = torch.randn_like(plain)
    gg = torch.autograd.grad(got, params, gy)
    wg = torch.autograd.grad(plain, params, gy)
    tol = 2e-5 if conv_mode == "fp32" else 1e-3
    assert T.rel(got, plain) <= tol
    for a, b, p in zip(gg, wg, params):
        assert T.rel(a, b) <= 5 * tol, tuple(p.shape)
