"""Size-independent properties at BASELINE.json's FULL sizes, where the CPU oracle would take minutes:
cfg 2 (blur / FusedLeakyReLU on (32,128,256,256)), cfg 3 (modulated 3x3 conv, batch 16, 512 channels, 64x64)
and the bit path at thousands of messages.  Each property holds for the reference's operators by definition
(linearity and adjointness of upfirdn2d / conv2d, scale invariance of the demodulated convolution,
encode -> decode round trip), so it pins the CUDA path at sizes the small-case oracle comparisons cannot reach.
Tolerances: exact-arithmetic properties 1e-5 relative; anything through the tf32 tensor path 1e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def _dot(a, b):
    return float((a.double() * b.double()).sum())


def _kernel(gain=1.0):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], device="cuda")
    k = torch.outer(k, k)
    return k / k.sum() * gain


def _nhwc(*shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)


# --------------------------------------------------------------------------------- cfg 2: upfirdn2d / blur
@pytest.mark.parametrize("pad", [(2, 2), (1, 1)])
def test_blur_cfg2_linearity_constant_and_adjoint(pad):
    from ideas_b200.stylegan2.op import upfirdn2d
    B, C, H = 32, 128, 256
    k = _kernel()
    x, y = _nhwc(B, C, H, H, seed=1), _nhwc(B, C, H, H, seed=2)
    bx, by = upfirdn2d(x, k, pad=pad), upfirdn2d(y, k, pad=pad)
    Ho = H + 2 * pad[0] - 3
    assert tuple(bx.shape) == (B, C, Ho, Ho)
    # linearity
    lin = upfirdn2d(0.75 * x - 2.0 * y, k, pad=pad)
    assert _rel(lin, 0.75 * bx - 2.0 * by) <= 1e-5
    # the taps sum to one: a constant image stays constant away from the zero padding
    const = upfirdn2d(torch.full_like(x, 3.0), k, pad=pad)
    inner = const[:, :, 3:-3, 3:-3]
    assert float((inner - 3.0).abs().max()) <= 1e-5
    # adjointness <blur(x), g> == <x, blur^T(g)>: checks the backward kernel at full size
    xr = x.clone().requires_grad_(True)
    out = upfirdn2d(xr, k, pad=pad)
    g = _nhwc(B, C, Ho, Ho, seed=3)
    (gx,) = torch.autograd.grad(out, xr, g)
    lhs, rhs = _dot(out.detach(), g), _dot(x, gx)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)


def test_resampling_pair_is_adjoint_full_size():
    """upfirdn2d(down=2) and upfirdn2d(up=2) with the flipped kernel are adjoint (the fused skip-branch kernels):
    <down(x), g> == <x, down^T(g)> on the Dreal first-block skip shape."""
    from ideas_b200.stylegan2.op import upfirdn2d
    B, C, H = 32, 64, 256
    k = _kernel()
    x = _nhwc(B, C, H, H, seed=4).requires_grad_(True)
    d = upfirdn2d(x, k, down=2, pad=(1, 1))
    assert tuple(d.shape) == (B, C, 128, 128)
    g = _nhwc(B, C, 128, 128, seed=5)
    (gx,) = torch.autograd.grad(d, x, g)
    lhs, rhs = _dot(d.detach(), g), _dot(x.detach(), gx)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    # the up-sampling skip of G: 1x1 conv at low resolution then upfirdn2d(up=2): energy-preserving taps (gain 4 = up^2)
    u = upfirdn2d(torch.ones(4, 8, 32, 32, device="cuda"), _kernel(4.0), up=2, pad=(2, 1))
    assert tuple(u.shape) == (4, 8, 64, 64) and float((u[:, :, 2:-2, 2:-2] - 1.0).abs().max()) <= 1e-5


def test_fused_leaky_relu_cfg2_definition_and_mask():
    from ideas_b200.stylegan2.op import fused_leaky_relu
    B, C, H = 32, 128, 256
    x = _nhwc(B, C, H, H, seed=6).requires_grad_(True)
    b = torch.randn(C, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)).requires_grad_(True)
    out = fused_leaky_relu(x, b)
    want = torch.nn.functional.leaky_relu(x.detach() + b.detach().view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert _rel(out, want) <= 1e-6
    g = _nhwc(B, C, H, H, seed=8)
    gx, gb = torch.autograd.grad(out, [x, b], g)
    mask = torch.where(want > 0, 1.0, 0.2) * 2 ** 0.5
    assert _rel(gx, g * mask) <= 1e-6
    assert _rel(gb, (g * mask).sum(dim=(0, 2, 3))) <= 1e-4       # 2M-term fp32 sums per channel


# --------------------------------------------------------------------------------- cfg 3: 3x3 conv, 512 -> 512
def test_conv_cfg3_adjoint_identities():
    """<conv(x, w), g> == <x, dgrad(g, w)> == <w, wgrad(x, g)> ties the three tensor-core GEMMs of the cfg-3
    layer to one another at full size (16 x 512 x 64 x 64, 512 -> 512, 3x3)."""
    from ideas_b200 import _lib as L
    from ideas_b200.stylegan2.op import conv as CV
    N, C, K, H = 16, 512, 512, 64
    x = _nhwc(N, C, H, H, seed=9).requires_grad_(True)
    wp = (torch.randn(9, K, C, device="cuda", generator=torch.Generator(device="cuda").manual_seed(10)) / (9 * C) ** 0.5).requires_grad_(True)
    geom = CV.Geom.forward(x.shape, K, 3, 3, 1, 1)
    y = CV.ConvFwd.apply(x, wp, None, geom, L.ACT_NONE, 0.2, 1.0)
    g = _nhwc(N, K, H, H, seed=11)
    gx, gw = torch.autograd.grad(y, [x, wp], g)
    a, b, c = _dot(y.detach(), g), _dot(x.detach(), gx), _dot(wp.detach(), gw)
    scale = float(y.detach().double().norm() * g.double().norm())
    assert abs(a - b) <= 1e-3 * scale and abs(a - c) <= 1e-3 * scale, (a, b, c, scale)
    # linearity in the input (tf32 rounding of the operands is not exactly linear: 1e-3 of the output scale)
    y2 = CV.ConvFwd.apply(2.0 * x.detach(), wp.detach(), None, geom, L.ACT_NONE, 0.2, 1.0)
    assert _rel(y2, 2.0 * y.detach()) <= 1e-6      # a power of two scales exactly


def test_modulated_conv_cfg3_demodulation_is_scale_invariant():
    """ModulatedConv2d with demodulation (stylegan2/model.py:239-248): multiplying every style by a > 0 leaves the
    output unchanged (the weights are renormalised per sample), and the per-sample output norm is pinned by the
    demodulation -- at the cfg-3 size."""
    from ideas_b200.stylegan2 import model as M
    torch.manual_seed(12)
    m = M.ModulatedConv2d(512, 512, 3, 2048).cuda()
    x = _nhwc(16, 512, 64, 64, seed=13)
    style = torch.rand(16, 2048, device="cuda", generator=torch.Generator(device="cuda").manual_seed(14)) * 2 - 1
    with torch.no_grad():
        y1 = m(x, style)
        m.modulation.weight.mul_(3.0)
        m.modulation.bias.mul_(3.0)                # s -> 3 s for every sample
        y2 = m(x, style)
    assert _rel(y2, y1) <= 1e-3
    assert tuple(y1.shape) == (16, 512, 64, 64) and bool(torch.isfinite(y1).all())


# --------------------------------------------------------------------------------- bit path at scale
@pytest.mark.parametrize("sigma", [1, 2, 4])
def test_bit_path_round_trip_large(sigma):
    """message -> secret tensor -> message (train.py:254-286) on 4096 messages of 256 groups.  With the jitter
    strictly inside the interval (delta = 0.45) the round trip is the identity and the XOR/popcount BER is exactly
    0.  At the reference's delta = 0.5 the jitter reaches the interval edge up to fp32 rounding (u = 1 - 2^-24
    lands ON the boundary for sigma = 4), so there the contract is bit-exactness with the oracle's decode of the
    same tensor, and a mismatch rate against the message that stays at the 1e-6 level."""
    from oracle import bits as OB
    from ideas_b200 import utils as U
    B, L = 4096, 256
    g = torch.Generator().manual_seed(15 + sigma)
    M = torch.randint(0, 2, (B, sigma * L), generator=g).float()
    Z = U.message_to_tensor(M, sigma, 0.45)
    assert tuple(Z.shape) == (B, L) and float(Z.abs().max()) <= 1.0
    back = U.tensor_to_message(Z, sigma)
    assert torch.equal(back.cpu(), M)
    pm, pb = U.pack_message(M.cuda()), U.decode_packed(Z, sigma)
    assert U.bit_error_rate(pm, pb, B * sigma * L) == 0.0
    flipped = M.clone()
    flipped[::7, 3] = 1 - flipped[::7, 3]
    assert int(U.bit_errors_packed(U.pack_message(flipped.cuda()), pb).item()) == len(range(0, B, 7))
    # the reference's own jitter width
    Z5 = U.message_to_tensor(M, sigma, 0.5)
    got = U.tensor_to_message(Z5, sigma).cpu().numpy()
    want = OB.tensor_to_message(Z5.cpu().numpy(), sigma)
    assert (got == want).all()
    assert float((got != M.numpy()).mean()) <= 1e-5


def test_reflect_pad_full_size_crop_and_adjoint():
    from ideas_b200.stylegan2.op.elementwise import reflect_pad
    x = _nhwc(32, 32, 256, 256, seed=20).requires_grad_(True)
    p = reflect_pad(x, 1)
    assert tuple(p.shape) == (32, 32, 258, 258)
    assert torch.equal(p[:, :, 1:-1, 1:-1], x.detach())
    assert torch.equal(p[:, :, 0, 1:-1], x.detach()[:, :, 1, :]) and torch.equal(p[:, :, 1:-1, -1], x.detach()[:, :, :, -2])
    g = _nhwc(32, 32, 258, 258, seed=21)
    (gx,) = torch.autograd.grad(p, x, g)
    lhs, rhs = _dot(p.detach(), g), _dot(x.detach(), gx)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
