"""Pin oracle/nets.py (functional networks + parameter spec + seeded init) to the reference:
state_dict contract and seeded-init hashes (contract.json), small-net outputs
(nets_small.pt), and the cfg-1 pipeline of SURVEY.md §8(d) (cfg1.pt)."""
import hashlib
import json
import os

import pytest
import torch

from oracle import nets as N
from oracle import bits as B


def sha_state(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def close(a, b, tol=2e-5):
    scale = b.abs().max().clamp_min(1e-6)
    return float((a - b).abs().max() / scale) <= tol


@pytest.fixture(scope="module")
def contract(golden_dir):
    return json.load(open(os.path.join(golden_dir, "contract.json")))


@pytest.mark.parametrize("cfgname", ["default256", "cfg1_64", "small", "N2_512"])
def test_state_dict_contract_and_seeded_init(contract, cfgname):
    entry = contract[cfgname]
    torch.manual_seed(0)
    for name, want in entry["nets"].items():          # same order as the reference created them
        spec = N.param_spec(name, **entry["cfg"])
        assert [[k, list(s)] for k, s, _ in spec] == want["keys"], name
        sd = N.init_state(name, **entry["cfg"])
        assert sha_state(sd) == want["sha256"], name


def test_small_nets_match_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "nets_small.pt"))
    sd = g["sd"]
    with torch.no_grad():
        S, T = N.encoder(sd["DisentanglementEncoder"], g["X"])
        assert close(S, g["S"]) and close(T, g["T"])
        assert close(N.structure_generator(sd["StructureGenerator"], g["Z"]), g["S2"])
        assert close(N.generator(sd["Generator"], g["S2"], g["T"]), g["img"])
        assert close(N.extractor(sd["TensorExtractor"], g["S"]), g["zhat"])
        dco, refin = N.cooccur_discriminator(sd["CooccurenceDiscriminator"], g["P"], g["Pref"], ref_batch=2)
        assert close(dco, g["dco"]) and close(refin, g["refin"])
        dco2, _ = N.cooccur_discriminator(sd["CooccurenceDiscriminator"], g["P"], ref_input=g["refin"])
        assert close(dco2, g["dco2"])
        assert close(N.distribution_discriminator(sd["DistributionDiscriminator"], g["T"]), g["ddist"])
        torch.manual_seed(g["dreal_seed"])
        dsd = N.init_state("ImageLevelDiscriminator", **g["cfg"])
        assert sha_state(dsd) == g["dreal_sha"]
        assert close(N.image_discriminator(dsd, g["X256"]), g["dreal"], 1e-4)


def test_cfg1_pipeline(golden_dir):
    g = torch.load(os.path.join(golden_dir, "cfg1.pt"))
    torch.manual_seed(0)
    cfg = dict(image_size=64)
    E, G = N.init_state("DisentanglementEncoder", **cfg), N.init_state("Generator", **cfg)
    Gs, Ex = N.init_state("StructureGenerator", **cfg), N.init_state("TensorExtractor", **cfg)
    assert sha_state(E) == g["sha"]["E"] and sha_state(G) == g["sha"]["G"]
    assert sha_state(Gs) == g["sha"]["Gstru"] and sha_state(Ex) == g["sha"]["Ex"]
    X = torch.rand(2, 3, 64, 64) * 2 - 1
    assert torch.equal(X, g["X"])
    with torch.no_grad():
        S1, T1 = N.encoder(E, X)
        Z = torch.rand(2, 1, 4, 4) * 2 - 1
        S2 = N.structure_generator(Gs, Z)
        Xh = N.generator(G, S2, T1)
        Sh, _ = N.encoder(E, Xh)
        Zh = N.extractor(Ex, Sh)
    for a, b in ((S1, g["S1"]), (T1, g["T1"]), (S2, g["S2"]), (Xh, g["Xh"]), (Sh, g["Sh"]), (Zh, g["Zh"])):
        assert close(a, b, 1e-4)
    hatM = B.tensor_to_message(g["Zh"].reshape(2, -1).numpy(), 1)
    assert (hatM == g["hatM"].numpy()).all()
