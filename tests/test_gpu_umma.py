"""Differential tests of the tcgen05/TMA convolution kernels (IDEAS_IMPL_UMMA) against the exact
fp32 FFMA kernels (IDEAS_IMPL_SIMT) of the same library, called straight through the C ABI on
identical device buffers, plus an fp64 torch reference for the TF32 error level.
Shapes follow the IDEAS networks at reduced batch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from ideas_b200 import _lib as L
    return L


def _p(t):
    from ideas_b200._tensor import ptr
    return ptr(t)


def _stream(t):
    from ideas_b200._tensor import stream_ptr
    return stream_ptr(t)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-9))


def conv_forward(x, wp, N, H, W, C, K, kh, kw, stride, pad, impl, out_scale=None, bias=None, act=0):
    L = _lib()
    OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    y = torch.full((N, OH, OW, K), float("nan"), device=x.device)
    L.call("ideas_conv2d_forward", _p(y), _p(x), _p(wp), _p(None), _p(out_scale), _p(bias), N, H, W, C, K, kh, kw, stride,
           pad, act, 0.2, 2 ** 0.5, impl, _stream(x))
    return y


def conv_dgrad(dy, wpt, N, H, W, C, K, kh, kw, stride, pad, OH, OW, impl, in_scale=None):
    L = _lib()
    dx = torch.full((N, H, W, C), float("nan"), device=dy.device)
    L.call("ideas_conv2d_dgrad", _p(dx), _p(dy), _p(wpt), _p(in_scale), _p(None), _p(None), N, H, W, C, K, kh, kw, stride,
           pad, OH, OW, 0, 0.2, 1.0, impl, _stream(dy))
    return dx


def conv_wgrad(x, dy, N, H, W, C, K, kh, kw, stride, pad, OH, OW, impl):
    L = _lib()
    dwp = torch.zeros((kh * kw, K, C), device=x.device)
    L.call("ideas_conv2d_wgrad", _p(dwp), _p(x), _p(dy), _p(None), _p(None), N, H, W, C, K, kh, kw, stride, pad, OH, OW,
           impl, _stream(x))
    return dwp


FWD_CASES = [
    # N, C, K, H, W, k, stride, pad        (role in IDEAS)
    (2, 32, 32, 16, 16, 3, 1, 1),          # minimal eligible
    (2, 64, 128, 16, 16, 3, 1, 1),         # G L0-L3 style, 16x16
    (3, 128, 256, 32, 32, 3, 1, 1),        # odd batch -> ragged sub-tile pairs
    (1, 512, 512, 64, 64, 3, 1, 1),        # cfg-3 layer (B=1)
    (2, 128, 128, 66, 66, 3, 1, 0),        # reflect-padded conv1 of E (pad 0 on H+2)
    (2, 64, 128, 65, 65, 3, 2, 0),         # Blur -> stride-2 conv (ResBlock down)
    (2, 256, 512, 17, 17, 3, 2, 0),        # E texture head 17 -> 8
    (2, 64, 128, 63, 63, 1, 2, 0),         # skip: Blur(1,1) -> 1x1 stride 2
    (2, 512, 512, 16, 16, 1, 1, 0),        # E structure head 1x1
    (4, 384, 384, 8, 8, 3, 1, 1),          # Dco mid layers, OC = 384 -> BLOCK_N 128
    (2, 96, 64, 20, 12, 3, 1, 1),          # non-square, C = 96
    (8, 384, 768, 4, 4, 3, 1, 1),          # 4x4 images: 8 images per 128-pixel box
    (1, 32, 64, 256, 256, 3, 1, 1),        # E 256x256 first block
]


@pytest.mark.parametrize("case", FWD_CASES)
def test_umma_forward_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    wp = torch.randn(k * k, K, C, device="cuda", generator=g) / (C * k * k) ** 0.5
    d = torch.rand(N, K, device="cuda", generator=g) + 0.5
    b = torch.randn(K, device="cuda", generator=g)
    ref = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_SIMT, d, b, 1)
    try:
        got = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_UMMA, d, b, 1)
    except RuntimeError as e:
        if "not eligible" in str(e):
            pytest.skip("below the tcgen05 size threshold: served by the FFMA kernel")
        raise
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)
    # plain (no epilogue)
    ref = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_SIMT)
    got = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_UMMA)
    assert rel(got, ref) <= 1e-3, rel(got, ref)


DGRAD_CASES = [
    # N, C(in of fwd), K(out of fwd), H, W (fwd input), k, stride, pad
    (2, 64, 128, 16, 16, 3, 1, 1),
    (2, 128, 128, 34, 34, 3, 1, 0),
    (2, 128, 64, 33, 33, 3, 2, 0),         # transposed conv 16 -> 33 (upsampling modconv) / dgrad of down conv
    (1, 256, 256, 65, 65, 3, 2, 0),
    (2, 64, 64, 31, 31, 1, 2, 0),          # k=1 transposed skip 16 -> 31
    (2, 512, 256, 17, 17, 3, 2, 0),
    (2, 32, 32, 129, 129, 3, 2, 0),
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_umma_dgrad_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(2)
    dy = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    wpt = torch.randn(k * k, C, K, device="cuda", generator=g) / (K * k * k) ** 0.5
    s = torch.rand(N, C, device="cuda", generator=g) + 0.5
    ref = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_SIMT, s)
    # AUTO: tcgen05 for every phase that has taps, FFMA for empty phases (k=1, stride 2)
    got = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_AUTO, s)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)


WGRAD_CASES = [
    (2, 32, 32, 16, 16, 3, 1, 1),
    (2, 64, 128, 16, 16, 3, 1, 1),
    (3, 128, 256, 32, 32, 3, 1, 1),
    (1, 512, 512, 64, 64, 3, 1, 1),
    (2, 128, 128, 66, 66, 3, 1, 0),
    (2, 64, 128, 65, 65, 3, 2, 0),
    (2, 64, 128, 63, 63, 1, 2, 0),
    (2, 512, 512, 16, 16, 1, 1, 0),
    (4, 384, 384, 8, 8, 3, 1, 1),
    (2, 96, 64, 20, 12, 3, 1, 1),
    (2, 128, 64, 33, 33, 3, 2, 0),
    (1, 32, 64, 256, 256, 3, 1, 1),        # 9-tap group (C = 32), 32x1 boxes
    (2, 128, 128, 130, 34, 3, 1, 0),       # filter-row groups, ragged boxes, pad 0
    (3, 64, 160, 21, 45, 3, 1, 1),         # odd sizes, K = 160
    (5, 256, 128, 9, 9, 3, 1, 1),          # 8x4 boxes hanging over 9x9 maps
    (2, 768, 384, 17, 17, 2, 1, 0),          # 2x2 valid conv
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_umma_wgrad_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    dy = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    ref = conv_wgrad(x, dy, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_SIMT)
    try:
        got = conv_wgrad(x, dy, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_UMMA)
    except RuntimeError as e:
        if "not eligible" in str(e):
            pytest.skip("tcgen05 wgrad not available for this shape")
        raise
    torch.cuda.synchronize()
    assert rel(got, ref) <= 1e-3, rel(got, ref)


# ---------------------------------------------------------------------------------------------
# halo-reuse kernel (conv_umma_halo_kernel): forced on ("halo" = 2) for every eligible launch and
# compared with the FFMA kernel; shapes cover ragged tiles in x and y, pad 0 / pad 1 tap sets, 2x2
# filters, channel counts that leave the 128-row weight tile partly empty, and the stride-2 data
# gradient whose four output phases use 1/2/2/4 taps with a strided destination.
class _halo:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        _lib().call("ideas_set_option", b"halo", self.mode)

    def __exit__(self, *a):
        _lib().call("ideas_set_option", b"halo", 1)


HALO_FWD_CASES = [
    # N, C, K, H, W, k, pad
    (2, 32, 128, 16, 16, 3, 1),
    (2, 64, 128, 16, 16, 3, 1),
    (3, 128, 256, 32, 32, 3, 1),
    (1, 512, 512, 64, 64, 3, 1),
    (2, 128, 128, 66, 66, 3, 0),
    (2, 96, 64, 20, 12, 3, 1),           # K = 64: half of the weight tile is TMA zero fill
    (1, 32, 64, 256, 256, 3, 1),
    (2, 64, 384, 37, 29, 3, 1),          # ragged in both directions, K = 3 x 128
    (1, 128, 128, 130, 258, 3, 0),       # wide rows
    (2, 64, 160, 9, 50, 3, 1),           # K = 160: second k-tile has 32 live rows
    (4, 384, 384, 14, 14, 3, 1),         # narrowest eligible rows (bwp = 16)
    (4, 768, 384, 17, 17, 2, 0),         # 2x2 valid conv (Dco head shape, larger image)
    (1, 64, 64, 64, 64, 3, 1),           # 32-wide tiles: TMA-store epilogue, K = 64 clipped by the unit
    (2, 32, 160, 32, 96, 3, 1),          # TMA-store epilogue, second k-tile has 32 live channels, 96-pixel rows
    (1, 128, 128, 37, 128, 3, 1),        # TMA-store epilogue, ragged in y only
]


@pytest.mark.parametrize("case", HALO_FWD_CASES)
def test_halo_forward_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, pad = case
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    wp = torch.randn(k * k, K, C, device="cuda", generator=g) / (C * k * k) ** 0.5
    d = torch.rand(N, K, device="cuda", generator=g) + 0.5
    b = torch.randn(K, device="cuda", generator=g)
    ref = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_SIMT, d, b, 1)
    n0 = L.launch_count()
    with _halo(2), _opt(b"pmh", 0, 1):
        got = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA, d, b, 1)
        plain = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA)
        n1 = L.launch_count()
        with _opt(b"halo_epi", 3, 1):        # transposed 128-bit-store epilogue everywhere (no TMA stores)
            transposed = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA, d, b, 1)
    torch.cuda.synchronize()
    assert n1 - n0 == 2
    assert torch.equal(got, transposed)      # the epilogue variants write the same bits
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)
    assert rel(plain, conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_SIMT)) <= 1e-3
    with _halo(0):
        old = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA, d, b, 1)
    # same tf32 products, different summation order only
    assert rel(got, old) <= 2e-5, rel(got, old)


@pytest.mark.parametrize("case", DGRAD_CASES + [(2, 128, 128, 64, 64, 3, 1, 1), (1, 256, 128, 40, 24, 3, 1, 1)])
def test_halo_dgrad_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(6)
    dy = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    wpt = torch.randn(k * k, C, K, device="cuda", generator=g) / (K * k * k) ** 0.5
    s = torch.rand(N, C, device="cuda", generator=g) + 0.5
    ref = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_SIMT, s)
    with _halo(2):
        got = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_AUTO, s)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)


PHASED_DGRAD_CASES = [c for c in DGRAD_CASES if c[6] == 2 and c[5] == 3] + [
    (2, 128, 256, 128, 128, 3, 2, 1),      # 64-pixel phase rows: 32-wide tiles, one strided TMA-store map per phase
    (2, 128, 256, 67, 65, 3, 2, 0),        # odd, non-square: the four phase grids differ in both directions
    (1, 64, 128, 130, 130, 3, 2, 1),       # padded stride-2 conv
    (2, 512, 512, 33, 33, 3, 2, 0),
    (3, 128, 128, 20, 36, 4, 2, 1),        # 4x4 filter: four taps in every phase
    (16, 128, 256, 17, 17, 3, 2, 0),       # more items than CTAs: the phase rotation over rounds
]


@pytest.mark.parametrize("case", PHASED_DGRAD_CASES)
def test_phased_dgrad_single_launch_matches_per_phase(case):
    """Option dgrad_phases: the stride^2 output phases of a strided data gradient in ONE halo launch (phase-
    interleaved items) give the same result as one launch per phase, and as the FFMA kernel."""
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(16)
    dy = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    wpt = torch.randn(k * k, C, K, device="cuda", generator=g) / (K * k * k) ** 0.5
    s = torch.rand(N, C, device="cuda", generator=g) + 0.5
    ref = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_SIMT, s)
    with _opt(b"dgrad_phases", 0, 1), _halo(2), _opt(b"pmh", 0, 1):
        per_phase = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_AUTO, s)
    with _opt(b"dgrad_phases", 2, 1), _halo(2):
        got = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_AUTO, s)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)
    assert rel(got, per_phase) <= 2e-5, rel(got, per_phase)


PAIR_FWD_CASES = [
    # N, C, K, H, W, k, stride, pad
    (2, 64, 256, 16, 16, 3, 1, 1),         # 4 sub-tiles: one cluster of two-sub-tile CTAs
    (3, 128, 256, 32, 32, 3, 1, 1),
    (1, 512, 512, 64, 64, 3, 1, 1),        # two 256-channel tiles
    (5, 96, 256, 20, 12, 3, 1, 1),         # ragged, odd number of sub-tiles (padding CTA)
    (2, 128, 512, 33, 33, 3, 2, 0),        # stride-2 source (per-parity maps)
    (7, 256, 256, 9, 9, 1, 1, 0),          # 1x1, 5 sub-tiles
    (16, 256, 256, 32, 32, 3, 1, 1),       # more CTAs than SMs: tail balancing with single-sub-tile clusters
    (4, 768, 256, 17, 17, 2, 1, 0),        # 2x2 valid conv
    (3, 64, 128, 32, 32, 3, 1, 1),         # 128-channel tiles (pair mode 2): each CTA stages 64 weight rows
    (2, 256, 384, 16, 16, 3, 1, 1),        # 384 = 3 x 128 (G's third block)
    (5, 96, 128, 20, 12, 3, 1, 1),
    (2, 128, 128, 65, 65, 3, 2, 0),
]


@pytest.mark.parametrize("case", PAIR_FWD_CASES)
def test_pair_forward_matches_simt(case):
    """conv_umma_fwd_pair_kernel (tcgen05 cta_group::2, option pair) against the FFMA kernel and the one-CTA kernel,
    with the fused epilogue (demodulation scale, bias, leaky ReLU)."""
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(26)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    wp = torch.randn(k * k, K, C, device="cuda", generator=g) / (C * k * k) ** 0.5
    d = torch.rand(N, K, device="cuda", generator=g) + 0.5
    b = torch.randn(K, device="cuda", generator=g)
    ref = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_SIMT, d, b, 1)
    with _opt(b"pmh", 0, 1), _halo(0):
        with _opt(b"pair", 0, 1):
            one = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_UMMA, d, b, 1)
        with _opt(b"pair", 2, 1):
            got = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_UMMA, d, b, 1)
            with _opt(b"pair_epi", 0, 1):
                direct = conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_UMMA, d, b, 1)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)
    assert rel(got, one) <= 2e-5, rel(got, one)
    assert torch.equal(got, direct)          # TMA-store epilogue vs direct stores: the same bits


@pytest.mark.parametrize("epi", [0, 1])
@pytest.mark.parametrize("case", [(1, 256, 256, 65, 65, 3, 2, 0), (2, 512, 256, 17, 17, 3, 2, 0), (2, 256, 128, 33, 33, 3, 1, 1),
                                  (3, 256, 64, 31, 31, 1, 2, 0), (2, 128, 256, 33, 29, 3, 2, 0)])
def test_pair_dgrad_matches_simt(case, epi):
    """Data gradients on the CTA-pair kernel (halo / pmh kernels off): strided destinations -- the TMA-store epilogue
    (pair_epi=1) writes each output phase through a tensor map whose strides skip the other phases -- and both
    epilogue variants."""
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(27)
    dy = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    wpt = torch.randn(k * k, C, K, device="cuda", generator=g) / (K * k * k) ** 0.5
    s = torch.rand(N, C, device="cuda", generator=g) + 0.5
    ref = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_SIMT, s)
    with _opt(b"pmh", 0, 1), _halo(0), _opt(b"dgrad_phases", 0, 1), _opt(b"pair", 2, 1), _opt(b"pair_epi", epi, 1):
        got = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_AUTO, s)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)


def test_tf32_error_level_vs_fp64():
    """The tensor path multiplies in TF32 (10-bit mantissa) and accumulates in fp32: report and bound
    its error against an fp64 convolution at the cfg-3 reduction length (K = 9 * 512)."""
    L = _lib()
    N, C, K, H = 1, 512, 64, 16
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(N, H, H, C, device="cuda", generator=g)
    w = torch.randn(K, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5
    wp = w.permute(2, 3, 0, 1).reshape(9, K, C).contiguous()
    want = F.conv2d(x.permute(0, 3, 1, 2).double().cpu(), w.double().cpu(), padding=1).permute(0, 2, 3, 1).cuda()
    e_simt = rel(conv_forward(x, wp, N, H, H, C, K, 3, 3, 1, 1, L.IMPL_SIMT), want)

    def signed_bias(got):
        return float(((got.double() - want) * want.sign()).mean() / want.abs().mean())

    got = conv_forward(x, wp, N, H, H, C, K, 3, 3, 1, 1, L.IMPL_UMMA)      # default: TFLOAT32 maps, TMA rounds
    e_round, b_round = rel(got, want), signed_bias(got)
    L.call("ideas_set_option", b"tma_tf32", 0)                              # FLOAT32 maps: tcgen05 truncates
    try:
        got = conv_forward(x, wp, N, H, H, C, K, 3, 3, 1, 1, L.IMPL_UMMA)
        e_trunc, b_trunc = rel(got, want), signed_bias(got)
    finally:
        L.call("ideas_set_option", b"tma_tf32", 1)
    print(f"max rel err vs fp64: FFMA fp32 {e_simt:.2e}; tcgen05 tf32 rounded by TMA {e_round:.2e} (bias {b_round:.1e}); "
          f"truncated {e_trunc:.2e} (bias {b_trunc:.1e})")
    assert e_simt <= 1e-5 and e_round <= 5e-4 and abs(b_round) <= 1e-4 and e_trunc <= 1.5e-3


# ---------------------------------------------------------------------------------------------
# pixel-major halo kernel (conv_umma_pmh_kernel): forced on ("pmh" = 2) for every eligible launch and compared with
# the FFMA kernel.  Shapes: the 32 / 64-channel layers it exists for (E / Dco / Dreal first blocks), output tiles of
# 32, 64 and 128 channels, ragged tiles in x and y, 2x2 filters, several k-tiles, the stride-2 data gradient with
# its strided destination, and the residual-merging epilogue (ideas_conv2d_forward_res).
class _opt:
    def __init__(self, name, mode, default=1):
        self.name, self.mode, self.default = name, mode, default

    def __enter__(self):
        _lib().call("ideas_set_option", self.name, self.mode)

    def __exit__(self, *a):
        _lib().call("ideas_set_option", self.name, self.default)


PMH_FWD_CASES = [
    # N, C, K, H, W, k, pad
    (2, 32, 64, 16, 16, 3, 1),
    (1, 32, 64, 258, 258, 3, 0),         # E first block (reflection-padded input)
    (3, 32, 64, 64, 64, 3, 1),           # Dco first block
    (1, 64, 128, 256, 256, 3, 1),        # Dreal first block (OCT = 128)
    (2, 64, 64, 37, 29, 3, 1),           # ragged both ways
    (2, 128, 64, 66, 66, 3, 0),          # data-gradient shape of a 64 -> 128 conv
    (2, 96, 32, 20, 12, 3, 1),           # OCT = 32
    (2, 64, 160, 9, 50, 3, 1),           # K = 160 = 5 x 32
    (2, 64, 192, 40, 24, 3, 1),          # K = 192 = 3 x 64
    (4, 384, 384, 14, 14, 3, 1),         # several channel steps, 3 k-tiles of 128
    (4, 768, 384, 17, 17, 2, 0),         # 2x2 valid conv
    (1, 128, 128, 130, 258, 3, 0),       # wide rows (bwp up to 256)
    (9, 64, 64, 5, 7, 3, 1),             # tiny maps
]


@pytest.mark.parametrize("case", PMH_FWD_CASES)
def test_pmh_forward_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, pad = case
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    wp = torch.randn(k * k, K, C, device="cuda", generator=g) / (C * k * k) ** 0.5
    d = torch.rand(N, K, device="cuda", generator=g) + 0.5
    b = torch.randn(K, device="cuda", generator=g)
    ref = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_SIMT, d, b, 1)
    n0 = L.launch_count()
    with _opt(b"pmh", 2):
        try:
            got = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA, d, b, 1)
        except RuntimeError as e:
            if "not eligible" in str(e):
                pytest.skip("below the tcgen05 size threshold: served by the FFMA kernel")
            raise
        plain = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA)
    torch.cuda.synchronize()
    assert L.launch_count() - n0 == 2
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)
    assert rel(plain, conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_SIMT)) <= 1e-3
    with _opt(b"pmh", 0), _opt(b"halo", 0):
        old = conv_forward(x, wp, N, H, W, C, K, k, k, 1, pad, L.IMPL_UMMA, d, b, 1)
    assert rel(got, old) <= 2e-5, rel(got, old)          # same tf32 products, different summation order only


@pytest.mark.parametrize("case", DGRAD_CASES + [(2, 128, 128, 64, 64, 3, 1, 1), (1, 64, 32, 40, 24, 3, 1, 1),
                                                (2, 64, 128, 130, 130, 3, 1, 0)])
def test_pmh_dgrad_matches_simt(case):
    L = _lib()
    N, C, K, H, W, k, stride, pad = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(8)
    dy = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    wpt = torch.randn(k * k, C, K, device="cuda", generator=g) / (K * k * k) ** 0.5
    s = torch.rand(N, C, device="cuda", generator=g) + 0.5
    ref = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_SIMT, s)
    with _opt(b"pmh", 2):
        got = conv_dgrad(dy, wpt, N, H, W, C, K, k, k, stride, pad, OH, OW, L.IMPL_AUTO, s)
    torch.cuda.synchronize()
    assert not torch.isnan(got).any(), "unwritten outputs"
    assert rel(got, ref) <= 1e-3, rel(got, ref)


@pytest.mark.parametrize("case", [(2, 64, 64, 33, 21, 3, 1, 1, 2), (2, 128, 128, 32, 32, 3, 1, 1, 2),
                                  (1, 512, 512, 32, 32, 3, 1, 1, 1), (2, 64, 128, 65, 65, 3, 2, 0, 1),
                                  (2, 24, 40, 9, 9, 3, 1, 1, 1)])
def test_conv_forward_residual_epilogue(case):
    """y = (lrelu(d * conv(x) + b) * gain + residual) * res_scale through ideas_conv2d_forward_res on the pixel-major
    halo kernel (pmh forced), the pixel-major kernel and the FFMA kernel, against the unfused composition."""
    L = _lib()
    N, C, K, H, W, k, stride, pad, pmh = case
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    wp = torch.randn(k * k, K, C, device="cuda", generator=g) / (C * k * k) ** 0.5
    d = torch.rand(N, K, device="cuda", generator=g) + 0.5
    b = torch.randn(K, device="cuda", generator=g)
    res = torch.randn(N, OH, OW, K, device="cuda", generator=g)
    rs = 2 ** -0.5
    want = (conv_forward(x, wp, N, H, W, C, K, k, k, stride, pad, L.IMPL_SIMT, d, b, 1) + res) * rs

    def run(impl):
        y = torch.full((N, OH, OW, K), float("nan"), device="cuda")
        L.call("ideas_conv2d_forward_res", _p(y), _p(x), _p(wp), _p(None), _p(d), _p(b), _p(res), rs, N, H, W, C, K, k, k,
               stride, pad, 1, 0.2, 2 ** 0.5, impl, _stream(x))
        return y

    assert rel(run(L.IMPL_SIMT), want) <= 1e-6
    if C % 32 == 0 and K % 32 == 0:
        with _opt(b"pmh", pmh):
            got = run(L.IMPL_UMMA)
        torch.cuda.synchronize()
        assert not torch.isnan(got).any()
        assert rel(got, want) <= 1e-3, rel(got, want)
