"""Pin oracle/bits.py to the reference's utils.py:74-97 (fixtures from the reference run,
tests/golden/bits.json) and to the RNG-free known answers of SURVEY.md App. E."""
import json
import os

import numpy as np
import pytest

from oracle import bits as B

KAT_Z = [-2, -1, -0.75, -0.5, -0.25, -1e-9, -6e-8, -5.96e-8, -1e-7, 0, 1e-9, 0.25, 0.5, 0.75, 1, 1.5]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return json.load(open(os.path.join(golden_dir, "bits.json")))


def test_app_e_literals():
    z = np.array([KAT_Z], dtype=np.float32)
    assert B.tensor_to_message(z, 1)[0].astype(int).tolist() == [0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1]
    s2 = "00 00 00 01 01 10 01 01 01 10 10 10 11 11 11 11".replace(" ", "")
    assert "".join(map(str, B.tensor_to_message(z, 2)[0].astype(int))) == s2
    s3 = "000 000 001 010 011 100 011 011 011 100 100 101 110 111 111 111".replace(" ", "")
    assert "".join(map(str, B.tensor_to_message(z, 3)[0].astype(int))) == s3
    M = np.array([[0, 0, 0, 1, 1, 0, 1, 1, 0, 1, 1, 0]], dtype=np.float32)
    assert B.message_to_tensor(M, 1, 0)[0].tolist() == [-.5, -.5, -.5, .5, .5, -.5, .5, .5, -.5, .5, .5, -.5]
    assert B.message_to_tensor(M, 2, 0)[0].tolist() == [-.75, -.25, .25, .75, -.25, .25]
    assert B.message_to_tensor(M, 3, 0)[0].tolist() == [-.875, .625, .625, .625]


def test_reference_fixtures(gold):
    z = np.array(gold["kat_z"], dtype=np.float32)
    for s, want in gold["kat_decode"].items():
        assert np.array_equal(B.tensor_to_message(z, int(s)), np.array(want, dtype=np.float32))
    M = np.array(gold["kat_m"], dtype=np.float32)
    for s, want in gold["kat_encode_delta0"].items():
        assert np.array_equal(B.message_to_tensor(M, int(s), 0), np.array(want, dtype=np.float32))
    for c in gold["random"]:
        M = np.array(c["M"], dtype=np.float32)
        Z = B.message_to_tensor(M, c["sigma"], c["delta"], np.array(c["u"], dtype=np.float32))
        assert np.array_equal(Z, np.array(c["Z"], dtype=np.float32)), c["sigma"]       # bit-exact fp32
        Mh = B.tensor_to_message(Z, c["sigma"])
        assert np.array_equal(Mh, np.array(c["Mh"], dtype=np.float32))
        assert B.ber(M, Mh) == c["ber"] == 0.0


def test_pack_roundtrip_and_popcount():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 256, 1000):
        a = rng.integers(0, 2, (3, n)).astype(np.float32)
        b = rng.integers(0, 2, (3, n)).astype(np.float32)
        pa, pb = B.pack_bits(a), B.pack_bits(b)
        assert np.array_equal(B.unpack_bits(pa, n), a)
        assert B.bit_errors_packed(pa, pb) == int(np.abs(a - b).sum())
        assert abs(B.bit_errors_packed(pa, pb) / a.size - B.ber(a, b)) < 1e-7
    assert B.pack_bits(np.zeros((2, 0))).shape == (2, 0)
