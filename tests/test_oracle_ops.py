"""Pin the op-level oracle (oracle/functional.py) to the reference's own CPU outputs
(tests/golden/ops.pt, produced by tests/golden/make_golden.py from /root/reference)."""
import os

import pytest
import torch

from oracle import functional as O


@pytest.fixture(scope="module")
def ops(golden_dir):
    return torch.load(os.path.join(golden_dir, "ops.pt"))


def close(a, b, tol=1e-5):
    scale = b.abs().max().clamp_min(1e-6)
    return float((a - b).abs().max() / scale) <= tol


def test_fused_leaky_relu(ops):
    for key in ("flrelu", "flrelu_2d"):
        c = ops[key]
        assert torch.equal(O.fused_leaky_relu(c["x"], c["bias"]), c["out"]), key


def test_bias_act_grad_variants():
    # fused_bias_act_kernel.cu:36-47: grad=1 masks by the sign of the saved *output*
    g = torch.Generator().manual_seed(0)
    x, b = torch.randn(2, 3, 4, 4, generator=g), torch.randn(3, generator=g)
    out = O.fused_leaky_relu(x, b)
    gy = torch.randn(2, 3, 4, 4, generator=g)
    gx = O.bias_act(gy, None, out, 3, 1, 0.2, 2 ** 0.5)
    xr = x.clone().requires_grad_(True)
    (O.fused_leaky_relu(xr, b) * gy).sum().backward()
    assert torch.allclose(gx, xr.grad, atol=0, rtol=1e-6)   # (g*a)*s vs (g*s)*a rounding
    assert torch.count_nonzero(O.bias_act(gy, None, out, 3, 2, 0.2, 1.0)) == 0
    assert torch.equal(O.bias_act(x, b, None, 1, 0, 0.2, 2.0), (x + b.view(1, 3, 1, 1)) * 2.0)


def test_upfirdn2d_all_modes(ops):
    for c in ops["upfirdn2d"]:
        got = O.upfirdn2d(c["x"], c["kernel"], c["up"], c["down"], tuple(c["pad"]))
        assert got.shape == c["out"].shape, c["name"]
        assert close(got, c["out"], 2e-6), c["name"]


def test_upfirdn2d_empty_and_identity():
    k1 = torch.ones(1, 1)
    x = torch.randn(1, 2, 5, 5)
    assert torch.equal(O.upfirdn2d(x, k1), x)
    assert O.upfirdn2d(torch.zeros(0, 2, 5, 5), O.make_kernel([1, 3, 3, 1]), pad=(2, 1)).shape == (0, 2, 5, 5)


def test_modulated_conv_forward_and_grads(ops):
    for c in ops["modconv"]:
        sd = c["sd"]
        x = c["x"].clone().requires_grad_(True)
        st = c["style"].clone().requires_grad_(True)
        w = sd["weight"].clone().requires_grad_(True)
        mw = sd["modulation.weight"].clone().requires_grad_(True)
        mb = sd["modulation.bias"].clone().requires_grad_(True)
        out = O.modulated_conv2d(x, st, w, mw, mb, demodulate=c["demodulate"], upsample=c["upsample"],
                                 downsample=c["downsample"], blur_kernel=sd.get("blur.kernel"))
        assert close(out, c["out"]), c["name"]
        grads = torch.autograd.grad((out * c["gy"]).sum(), [x, st, w, mw, mb])
        for gi, (a, b) in enumerate(zip(grads, c["grads"])):
            assert close(a, b, 2e-5), (c["name"], gi)


def test_styled_conv_with_noise_and_torgb(ops):
    c = ops["styledconv_noise"]
    sd = c["sd"]
    out = O.styled_conv(c["x"], c["style"], sd["conv.weight"], sd["conv.modulation.weight"],
                        sd["conv.modulation.bias"], sd["activate.bias"], noise=c["noise"],
                        noise_weight=sd["noise.weight"])
    assert close(out, c["out"])
    c = ops["torgb"]
    sd = c["sd"]
    y = O.modulated_conv2d(c["x"], c["style"], sd["conv.weight"], sd["conv.modulation.weight"],
                           sd["conv.modulation.bias"], demodulate=False) + sd["bias"]
    y = y + O.upfirdn2d(c["skip"], sd["upsample.kernel"], up=2, pad=(2, 1))   # Upsample, model.py:33-51
    assert close(y, c["out"])
