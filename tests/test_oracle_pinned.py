"""CPU, build container only: the oracle restatement (oracle/m3gnet_oracle.py) against the LIVE reference imported
from /root/reference through the shims of oracle/live_reference.py — beyond the committed fixtures: fresh seeds,
several hyper-parameter sets, the triplet enumeration and the basis operators.  Skipped where /root/reference is
absent (the GPU box); the fixtures under tests/golden/ (tests/test_oracle_golden.py) are what travels."""
import numpy as np
import pytest
import torch

from oracle import live_reference as lr
from oracle import m3gnet_oracle as O

pytestmark = pytest.mark.skipif(not lr.available(), reason="/root/reference is not present")

KEYS = ["edge_distances", "triplet_angles", "edge_weights", "x", "edge_attr", "scaled_atomic_energies",
        "scaled_total_energy", "total_energy", "forces", "stresses"]


def _close(a, b, rtol, atol):
    a, b = a.detach().double(), b.detach().double()
    scale = b.abs().max().item() if b.numel() else 0.0
    return (a - b).abs().max().item() <= atol + rtol * scale if a.numel() else True


@pytest.mark.parametrize("l_max,n_max,dim,blocks,seed", [(3, 3, 64, 3, 0), (2, 4, 32, 2, 1), (4, 3, 16, 1, 2)])
def test_whole_model_against_the_live_reference(l_max, n_max, dim, blocks, seed):
    build_model, compute_threebody, _ = lr.import_reference()
    torch.set_num_threads(1)
    lat, cart, z = O.mpf_like_structure(3 + seed)
    g1 = O.build_graph(lat, cart, z, 5.0, 4.0)
    g2 = O.build_graph(*O.fcc_supercell(2, jitter=0.1, seed=seed), 5.0, 4.0)
    for g, n in ((g1, len(cart)), (g2, 32)):
        tri, nti, ntij = compute_threebody(n, g["edge_index"], g["edge_distances_build"], 4.0)
        assert torch.equal(tri, g["triplet_edge_index"]) and torch.equal(nti, g["num_triplet_i"])
        assert torch.equal(ntij, g["num_triplet_ij"])
    b = O.collate([g1, g2])
    torch.manual_seed(seed)
    model = build_model(5.0, 4.0, l_max, n_max, 95, dim, blocks)
    sd = {k: (v.detach() * 2 if k.endswith("weight") else v.detach().clone()) for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    fac = torch.rand(l_max, n_max, generator=torch.Generator().manual_seed(seed)) + 0.5
    for m in model.model:
        if hasattr(m, "nsb"):
            m.nsb.factors = fac
    ref = model(lr.as_reference_graph(b))
    hp = O.HyperParams(l_max=l_max, n_max=n_max, embedding_dim=dim, num_blocks=blocks)
    out = O.forward(sd, hp, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()}, factors=fac)
    for k in KEYS:
        assert _close(out[k], ref[k], 5e-6, 2e-8), k


def test_noise_valued_bessel_factors_are_reproduced_bitwise():
    """Quirk Q1: the reference's normalisation table is fp32 round-off of the host's sin / cos."""
    _, _, inter = lr.import_reference()
    for l_max, n_max, rc in ((3, 3, 5.0), (9, 10, 4.2)):
        ref = inter.NormalizedSphericalBessel(cutoff=rc, l_max=l_max, n_max=n_max).factors
        assert torch.equal(O.bessel_factors(rc, l_max, n_max), ref)


def test_basis_operators_against_the_live_reference():
    _, _, inter = lr.import_reference()
    x = torch.linspace(1e-3, 40.0, 257, requires_grad=True)
    c = torch.linspace(-1.0, 1.0, 101, requires_grad=True)
    for l in range(9):
        for mine, theirs, arg in ((O.spherical_bessel, inter.spherical_bessel, x),
                                  (O.legendre_cos, inter.legendre_cos, c)):
            ya, yb = mine(arg, l), theirs(arg, l)
            go = torch.full_like(ya, 0.7)
            (ga,) = torch.autograd.grad(ya, arg, grad_outputs=go)
            (gb,) = torch.autograd.grad(yb, arg, grad_outputs=go)
            assert torch.equal(ya, yb) and torch.equal(ga, gb), (mine.__name__ if hasattr(mine, "__name__") else l, l)
    r = torch.linspace(0.1, 6.0, 60)
    assert torch.equal(O.cutoff_function(r, 4.0), inter.cutoff_function(r, 4.0))
