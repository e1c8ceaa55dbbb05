"""GPU: the whole drop-in model (build_model → model(batch)) against the live-reference fixtures, plus the
reference test-suite's properties (tests/test_model.py, tests/test_invariance.py) restated."""
import numpy as np
import pytest
import torch

from oracle import m3gnet_oracle as O
from tests.util import golden, graph_dict, report, state_dict_of, to_batch

pytestmark = pytest.mark.gpu

INTERMEDIATE = ["edge_distances", "triplet_angles", "edge_weights", "x", "edge_attr", "scaled_atomic_energies",
                "scaled_total_energy"]


def _energy_force_parity(out, g, prefix, n_atoms):
    """Parity protocol of SURVEY §8(c): absolute AND relative bounds (tolerances from BASELINE.json north_star:
    1e-5 eV/atom, 1e-4 eV/Å in fp32)."""
    E = out["total_energy"].detach().cpu().double()
    E_ref = torch.from_numpy(g[prefix + "total_energy"]).double()
    F = out["forces"].detach().cpu().double()
    F_ref = torch.from_numpy(g[prefix + "forces"]).double()
    dE = (E - E_ref).abs().max().item() / n_atoms
    dF = (F - F_ref).abs().max().item()
    Fmax = F_ref.abs().max().item()
    print(f"[parity] {prefix} |dE|/atom={dE:.3e} eV  max|dF|={dF:.3e} eV/A  max|F|={Fmax:.3e}  "
          f"rel dF={dF / Fmax:.3e}  sumF={F.sum(0).abs().max().item():.2e}")
    assert dE <= 1e-5 and dE <= 1e-5 * max(E_ref.abs().max().item() / n_atoms, 1e-3) + 1e-7
    assert dF <= 1e-4 and dF <= 1e-3 * Fmax
    report(prefix + "stresses", out["stresses"], g[prefix + "stresses"], 1e-8, 2e-3)


def _default_model(device, sd):
    from torch_m3gnet_b200 import build_model

    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    model.load_state_dict(sd)
    return model


def test_c1_default_model_energy_forces(device):
    g = golden("c1_default")
    sd = state_dict_of(g)
    model = _default_model(device, sd)
    fac = torch.from_numpy(g["factors"]).to(device)
    for i in (6, 8, 10):
        model.model[i].nsb.factors = fac  # the live reference's (noise-valued, host-dependent) table
    b = to_batch(graph_dict(g), device)
    out = model(b)
    assert out is b and not b["pos"].requires_grad
    for k in INTERMEDIATE:
        report("c1." + k, out[k], g["out." + k], 1e-6, 5e-6)
    _energy_force_parity(out, g, "out.", 32)
    # amplified weights (x3): forces O(0.1) eV/A make the absolute tolerance meaningful
    model.load_state_dict({k: (v * 3 if k.endswith("weight") else v) for k, v in sd.items()})
    b = to_batch(graph_dict(g), device)
    out = model(b)
    for k in INTERMEDIATE:
        report("c1x3." + k, out[k], g["out3." + k], 1e-5, 1e-5)
    _energy_force_parity(out, g, "out3.", 32)


def test_tio2_two_species(device):
    g, c1 = golden("tio2_default"), golden("c1_default")
    sd = state_dict_of(c1)
    model = _default_model(device, sd)
    out = model(to_batch(graph_dict(g), device))
    _energy_force_parity(out, g, "out.", 32)
    model.load_state_dict({k: (v * 3 if k.endswith("weight") else v) for k, v in sd.items()})
    out = model(to_batch(graph_dict(g), device))
    _energy_force_parity(out, g, "out3.", 32)


def _small_model(device, g):
    from torch_m3gnet_b200 import build_model

    rc = float(g["cutoff"])
    model = build_model(rc, rc, 2, 3, 93, 17, 2, device=device)
    model.load_state_dict(state_dict_of(g))
    return model


def test_small_batch_reference_test_config(device):
    """embedding_dim=17, l_max=2, two structures in one batch (reference tests/conftest.py:150-178)."""
    g = golden("small_batch")
    model = _small_model(device, g)
    out = model(to_batch(graph_dict(g), device))
    for k in INTERMEDIATE:
        report("small." + k, out[k], g["out." + k], 1e-6, 5e-6)
    _energy_force_parity(out, g, "out.", 6)
    assert not torch.isnan(out["x"]).any() and not torch.isnan(out["edge_attr"]).any()


def test_triplet_permutation_invariance(device):
    """reference tests/test_model.py:21-38 — in-place permutation of the triplet list, same batch object."""
    g = golden("small_batch")
    model = _small_model(device, g)
    b = to_batch(graph_dict(g), device)
    e1 = model(b)["edge_attr"].clone()
    f1 = b["forces"].clone()
    T = b["triplet_edge_index"].size(1)
    perm = torch.randperm(T, device=device)
    b["triplet_edge_index"][0] = b["triplet_edge_index"][0][perm]
    b["triplet_edge_index"][1] = b["triplet_edge_index"][1][perm]
    e2 = model(b)["edge_attr"].clone()
    torch.testing.assert_close(e1, e2)
    torch.testing.assert_close(f1, b["forces"])
    assert torch.equal(e1, e2), "canonical CSR makes the result independent of the triplet order bit for bit"


def test_batch_order(device):
    """reference tests/test_model.py:59-78 — batched energies equal per-graph energies."""
    from torch_m3gnet_b200.data.material_graph import Batch

    g = golden("small_batch")
    model = _small_model(device, g)
    gd = graph_dict(g)
    out = model(to_batch(gd, device))
    # split the fixture batch back into its two graphs
    energies = []
    n0 = int((gd["batch"] == 0).sum())
    e0 = int((gd["edge_index"][0] < n0).sum())
    t0 = int((gd["triplet_edge_index"][0] < e0).sum())
    parts = [
        dict(pos=gd["pos"][:n0], atom_types=gd["atom_types"][:n0], num_triplet_i=gd["num_triplet_i"][:n0],
             edge_index=gd["edge_index"][:, :e0], edge_cell_shift=gd["edge_cell_shift"][:e0],
             num_triplet_ij=gd["num_triplet_ij"][:e0], triplet_edge_index=gd["triplet_edge_index"][:, :t0],
             lattice=gd["lattice"][0:1], batch=torch.zeros(n0, dtype=torch.long)),
        dict(pos=gd["pos"][n0:], atom_types=gd["atom_types"][n0:], num_triplet_i=gd["num_triplet_i"][n0:],
             edge_index=gd["edge_index"][:, e0:] - n0, edge_cell_shift=gd["edge_cell_shift"][e0:],
             num_triplet_ij=gd["num_triplet_ij"][e0:], triplet_edge_index=gd["triplet_edge_index"][:, t0:] - e0,
             lattice=gd["lattice"][1:2], batch=torch.zeros(gd["pos"].shape[0] - n0, dtype=torch.long)),
    ]
    for p in parts:
        energies.append(model(to_batch(p, device))["total_energy"])
    torch.testing.assert_close(out["total_energy"], torch.cat(energies))


def test_forces_vs_finite_differences(device):
    """reference tests/test_model.py:90-120 (delta 1e-2, atol 1e-3, rtol 1e-2), amplified weights."""
    g = golden("small_batch")
    model = _small_model(device, g)
    sd = state_dict_of(g)
    model.load_state_dict({k: (v * 2 if k.endswith("weight") else v) for k, v in sd.items()})
    gd = graph_dict(g)
    b = to_batch(gd, device)
    forces = model(b)["forces"].clone()
    delta = 1e-2
    for node in range(b["num_nodes"]):
        bidx = int(b["batch"][node])
        for d in range(3):
            bp = b.clone()
            bp["pos"][node, d] += delta
            ep = model(bp)["total_energy"][bidx]
            bm = b.clone()
            bm["pos"][node, d] -= delta
            em = model(bm)["total_energy"][bidx]
            torch.testing.assert_close(forces[node, d], -(ep - em) / (2 * delta), atol=1e-3, rtol=1e-2)


def test_three_body_path_is_exercised_end_to_end(device):
    """With an O(1) factor table the three-body term changes energies and forces (quirk Q1 hides it otherwise):
    compare the whole model against the oracle run with the same injected table."""
    g = golden("c1_default")
    sd = {k: (v * 2 if k.endswith("weight") else v) for k, v in state_dict_of(g).items()}
    torch.manual_seed(11)
    fac = torch.rand(3, 3) + 0.5
    model = _default_model(device, sd)
    for i in (6, 8, 10):
        model.model[i].nsb.factors = fac.to(device)
    gd = graph_dict(g)
    out = model(to_batch(gd, device))
    ref = O.forward(sd, O.HyperParams(), {k: v.clone() for k, v in gd.items()}, factors=fac, create_graph=False)
    ref0 = O.forward(sd, O.HyperParams(), {k: v.clone() for k, v in gd.items()}, create_graph=False)
    moved = (ref["forces"] - ref0["forces"]).abs().max().item()
    print(f"[parity] three-body contribution to forces with O(1) factors: {moved:.3e} eV/A")
    assert moved > 1e-4
    report("tbmodel.energy", out["total_energy"], ref["total_energy"], 32 * 1e-5, 0)
    report("tbmodel.forces", out["forces"], ref["forces"], 1e-4, 1e-3)
    report("tbmodel.edge_attr", out["edge_attr"], ref["edge_attr"], 2e-5, 2e-5)


def test_length_and_energy_scale(device):
    from torch_m3gnet_b200 import build_model

    g = golden("c1_default")
    sd = state_dict_of(g)
    elem = torch.linspace(-1, 1, 95)
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, elemental_energies=elem, energy_scale=2.5, length_scale=1.25,
                        device=device)
    model.load_state_dict(sd)
    gd = graph_dict(g)
    out = model(to_batch(gd, device))
    hp = O.HyperParams(energy_scale=2.5, length_scale=1.25, elemental_energies=elem)
    ref = O.forward(sd, hp, {k: v.clone() for k, v in gd.items()}, create_graph=False)
    report("scaled.energy", out["total_energy"], ref["total_energy"], 32 * 1e-5, 0)
    report("scaled.forces", out["forces"], ref["forces"], 1e-6, 1e-3)
    report("scaled.elemental", out["elemental_energies"], ref["elemental_energies"], 0, 0)


def test_cuda_graph_replay_matches_eager(device):
    """GraphedStep (CUDA-graph replay of forward + adjoint kernels on a fixed graph) reproduces the eager call, also
    after the positions were changed in the static buffer (neighbour list kept, as between Verlet rebuilds)."""
    from torch_m3gnet_b200.graphed import GraphedStep

    g = golden("c1_default")
    sd = {k: (v * 3 if k.endswith("weight") else v) for k, v in state_dict_of(g).items()}
    model = _default_model(device, sd)
    b = to_batch(graph_dict(g), device)
    step = GraphedStep(model, b)
    out = step()
    ref = model(to_batch(graph_dict(g), device))
    assert torch.equal(out["total_energy"], ref["total_energy"]) and torch.equal(out["forces"], ref["forces"])
    torch.manual_seed(1)
    new_pos = ref["pos"] + 0.01 * torch.randn_like(ref["pos"])
    out = step(pos=new_pos)
    b2 = to_batch(graph_dict(g), device)
    b2["pos"] = new_pos.clone()
    ref2 = model(b2)
    assert not torch.equal(ref2["forces"], ref["forces"])
    report("graphed.E", out["total_energy"], ref2["total_energy"], 1e-6, 1e-6)
    report("graphed.F", out["forces"], ref2["forces"], 1e-6, 1e-5)


@pytest.mark.parametrize("l_max,n_max,dim,blocks", [(3, 4, 64, 2), (4, 4, 64, 1), (2, 2, 64, 2), (3, 3, 128, 1),
                                                    (1, 1, 32, 3), (5, 3, 64, 1), (9, 10, 64, 1)])
def test_off_default_hyper_parameters(device, l_max, n_max, dim, blocks):
    """Other (l_max, n_max, width, depth): n_max = 4 takes the tensor-core forward with the FMA backward, l_max = 4
    and width != 64 the generic kernels; whole model against the oracle on a two-species cell with O(1) factors."""
    from torch_m3gnet_b200 import build_model

    hp = O.HyperParams(l_max=l_max, n_max=n_max, embedding_dim=dim, num_blocks=blocks)
    sd = O.init_params(hp, seed=7, gain=2.0)
    lat, cart, z = O.mpf_like_structure(2)
    gd = O.collate([O.build_graph(lat, cart, z, 5.0, 4.0), O.build_graph(*O.fcc_supercell(2, jitter=0.1, seed=4), 5.0, 4.0)])
    fac = torch.rand(l_max, n_max, generator=torch.Generator().manual_seed(1)) + 0.5
    model = build_model(5.0, 4.0, l_max, n_max, 95, dim, blocks, device=device)
    model.load_state_dict(sd)
    for m in model.model:
        if hasattr(m, "nsb"):
            m.nsb.factors = fac.to(device)
    out = model(to_batch(gd, device))
    ref = O.forward(sd, hp, {k: v.clone() for k, v in gd.items()}, factors=fac, create_graph=False)
    n = gd["pos"].shape[0]
    report("hp.energy", out["total_energy"], ref["total_energy"], n * 1e-5, 1e-5)
    report("hp.forces", out["forces"], ref["forces"], 1e-4, 1e-3)
    report("hp.x", out["x"], ref["x"], 2e-5, 2e-5)


def test_degenerate_graphs(device):
    """An isolated atom (no bonds), a dimer (bonds, no triplets) and a normal cell in one batch; and a batch whose only
    structure has no bond at all."""
    from torch_m3gnet_b200.data.material_graph import Batch
    from torch_m3gnet_b200.data.structure import Structure

    big = 30.0 * np.eye(3)
    structs = [(big, np.array([[1.0, 2.0, 3.0]]), np.array([6])),
               (big, np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 2.2]]), np.array([8, 1])),
               O.fcc_supercell(2, jitter=0.05, seed=2)]
    g = golden("c1_default")
    sd = {k: (v * 2 if k.endswith("weight") else v) for k, v in state_dict_of(g).items()}
    model = _default_model(device, sd)
    b = Batch.from_structures([Structure(l, [int(v) for v in zz], c, coords_are_cartesian=True) for l, c, zz in structs],
                              5.0, 4.0, device=device)
    out = model(b)
    gd = O.collate([O.build_graph(l, c, zz, 5.0, 4.0) for l, c, zz in structs])
    assert torch.equal(b["edge_index"].cpu(), gd["edge_index"])
    ref = O.forward(sd, O.HyperParams(), {k: v.clone() for k, v in gd.items()}, create_graph=False)
    report("degenerate.energy", out["total_energy"], ref["total_energy"], 35 * 1e-5, 1e-5)
    report("degenerate.forces", out["forces"], ref["forces"], 1e-4, 1e-3)
    assert out["forces"][0].abs().max().item() == 0.0
    lone = Batch.from_structures([Structure(big, [6], np.array([[1.0, 2.0, 3.0]]), coords_are_cartesian=True)], 5.0, 4.0,
                                 device=device)
    out1 = model(lone)
    assert lone["edge_index"].shape[1] == 0 and torch.isfinite(out1["total_energy"]).all()
    assert out1["forces"].abs().max().item() == 0.0
