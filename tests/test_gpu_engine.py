"""GPU: the whole-step C executor (m3g_step_run, torch_m3gnet_b200/engine.py) against the per-operator
torch.autograd.Function path it replaces for the default model shape, and against the live-reference fixtures."""
import numpy as np
import pytest
import torch

from oracle import m3gnet_oracle as O
from tests.util import golden, graph_dict, report, state_dict_of, to_batch

pytestmark = pytest.mark.gpu

KEYS = ["scaled_pos", "scaled_lattice", "elemental_energies", "edge_distances", "triplet_angles", "edge_weights", "x",
        "edge_attr", "scaled_atomic_energies", "scaled_total_energy", "total_energy", "forces", "stresses"]


def _model(device, sd, blocks=3, **kw):
    from torch_m3gnet_b200 import build_model

    model = build_model(5.0, 4.0, 3, 3, 95, 64, blocks, device=device, **kw)
    model.load_state_dict(sd)
    return model


def _both_paths(model, make_batch):
    from torch_m3gnet_b200 import engine
    from torch_m3gnet_b200.data.material_graph import get_plan

    res = {}
    old = engine.ENABLED
    try:
        for flag in (True, False):
            engine.ENABLED = flag
            b = make_batch()
            assert model.step_engine().supports(b, get_plan(b)) == flag
            out = model(b)
            res[flag] = {k: (out[k].clone() if torch.is_tensor(out[k]) else None) for k in KEYS}
            assert not out["pos"].requires_grad
    finally:
        engine.ENABLED = old
    return res[True], res[False]


@pytest.mark.parametrize("fixture,amp", [("c1_default", 3.0), ("tio2_default", 3.0)])
def test_engine_equals_operator_path_and_reference(device, fixture, amp):
    g, c1 = golden(fixture), golden("c1_default")
    sd = {k: (v * amp if k.endswith("weight") else v) for k, v in state_dict_of(c1).items()}
    model = _model(device, sd)
    fac = torch.rand(3, 3, generator=torch.Generator().manual_seed(2)) + 0.5
    for i in (6, 8, 10):
        model.model[i].nsb.factors = fac.to(device)   # O(1) table: the three-body adjoint is visible in the forces
    gd = graph_dict(g)
    eng, ops = _both_paths(model, lambda: to_batch(gd, device))
    for k in KEYS:
        if ops[k] is None:
            assert eng[k] is None
            continue
        scale = ops[k].abs().max().item()
        d = (eng[k] - ops[k]).abs().max().item()
        print(f"[engine] {fixture} {k}: max|ref|={scale:.3e} max|diff|={d:.3e}")
        assert d <= 2e-6 * scale + 1e-9, k
    ref = O.forward(sd, O.HyperParams(), {k: v.clone() for k, v in gd.items()}, factors=fac, create_graph=False)
    n = gd["pos"].shape[0]
    report("engine.energy", eng["total_energy"], ref["total_energy"], n * 1e-5, 1e-5)
    report("engine.forces", eng["forces"], ref["forces"], 1e-4, 1e-3)
    report("engine.stresses", eng["stresses"], ref["stresses"], 1e-5, 1e-3)


def test_engine_on_the_live_reference_fixture_with_default_factors(device):
    """Unmodified default model (noise-valued Bessel factors, quirk Q1) on config 1: the fixture's own outputs."""
    g = golden("c1_default")
    model = _model(device, state_dict_of(g))
    b = to_batch(graph_dict(g), device)
    from torch_m3gnet_b200.data.material_graph import get_plan

    assert model.step_engine().supports(b, get_plan(b))
    out = model(b)
    report("engine.c1.energy", out["total_energy"], g["out.total_energy"], 32 * 1e-5, 1e-5)
    report("engine.c1.forces", out["forces"], g["out.forces"], 1e-4, 1e-3)
    for k in ("edge_distances", "triplet_angles", "edge_weights", "x", "edge_attr", "scaled_atomic_energies", "stresses"):
        report("engine.c1." + k, out[k], g["out." + k], 2e-5, 2e-5)


def test_engine_ragged_batch_scales_and_elemental_energies(device):
    """Ragged multi-species batch from the GPU builder, length / energy scales and a non-zero elemental table."""
    from torch_m3gnet_b200 import Batch

    hp = O.HyperParams(energy_scale=2.5, length_scale=1.3, elemental_energies=torch.linspace(-1.0, 1.0, 95))
    sd = O.init_params(hp, seed=3, gain=2.0)
    model = _model(device, sd, energy_scale=2.5, length_scale=1.3, elemental_energies=hp.elemental_energies)
    structs = [O.mpf_like_structure(s) for s in (1, 2, 5)] + [O.fcc_supercell(2, jitter=0.1, seed=9)]
    lat = np.stack([s[0] for s in structs])
    cart = np.concatenate([s[1] for s in structs])
    z = np.concatenate([s[2] for s in structs])
    sizes = [len(s[1]) for s in structs]

    def make():
        return Batch.from_arrays(lat, cart, z, sizes, 5.0, 4.0, device=device)

    eng, ops = _both_paths(model, make)
    for k in KEYS:
        scale = ops[k].abs().max().item()
        assert (eng[k] - ops[k]).abs().max().item() <= 2e-6 * scale + 1e-9, k
    gd = O.collate([O.build_graph(s[0], s[1], s[2], 5.0, 4.0) for s in structs])
    ref = O.forward(sd, hp, {k: v.clone() for k, v in gd.items()}, create_graph=False)
    report("engine.ragged.energy", eng["total_energy"], ref["total_energy"], max(sizes) * 1e-5, 1e-5)
    report("engine.ragged.forces", eng["forces"], ref["forces"], 1e-4, 1e-3)


def test_engine_with_permuted_and_incomplete_triplet_lists(device):
    """A randomly permuted triplet list (reference tests/test_model.py:26-34) is canonicalised by the plan and still
    takes the executor (cos output in the caller's order); an incomplete (hand-thinned) list, other widths and
    keep_graph stay on the operator path."""
    from torch_m3gnet_b200 import build_model
    from torch_m3gnet_b200.data.material_graph import get_plan

    g = golden("c1_default")
    model = _model(device, state_dict_of(g))
    b = to_batch(graph_dict(g), device)
    out0 = model(b)
    e0, f0, cos0 = out0["total_energy"].clone(), out0["forces"].clone(), out0["triplet_angles"].clone()
    gd = graph_dict(g)
    perm = torch.randperm(gd["triplet_edge_index"].shape[1], generator=torch.Generator().manual_seed(0))
    gd["triplet_edge_index"] = gd["triplet_edge_index"][:, perm]
    bp = to_batch(gd, device)
    assert model.step_engine().supports(bp, get_plan(bp))
    outp = model(bp)
    assert torch.equal(outp["total_energy"], e0) and torch.equal(outp["forces"], f0)
    assert torch.equal(outp["triplet_angles"], cos0[perm.to(device)])
    gd["triplet_edge_index"] = gd["triplet_edge_index"][:, : perm.numel() // 2]  # not the full pair matrix any more
    bh = to_batch(gd, device)
    assert not model.step_engine().supports(bh, get_plan(bh))
    ref = O.forward(state_dict_of(g), O.HyperParams(), {k: v.clone() for k, v in gd.items()}, create_graph=False)
    report("thinned.energy", model(bh)["total_energy"], ref["total_energy"], 32 * 1e-5, 1e-5)
    wide = build_model(5.0, 4.0, 3, 3, 95, 32, 1, device=device)
    assert not wide.step_engine().ok
    model.keep_graph = True
    outk = model(to_batch(graph_dict(g), device))
    model.keep_graph = False
    assert outk["total_energy"].grad_fn is not None
    torch.testing.assert_close(outk["forces"], f0, rtol=1e-5, atol=1e-9)
