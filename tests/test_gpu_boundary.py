"""GPU: the drop-in boundary at the reference's full range: basis sizes up to l_max = 9 / n_max = 10,
GatedMLP.forward on its own, ThreeBodyInteration from public keys, bonds not grouped by source atom, loud failures for
out-of-range atom types / double backward / stale CUDA-graph captures, kernels following the tensors' device."""
import numpy as np
import pytest
import torch

from oracle import m3gnet_oracle as O
from tests.util import golden, graph_dict, report, state_dict_of, to_batch

pytestmark = pytest.mark.gpu

OUT_KEYS = ["edge_distances", "triplet_angles", "edge_weights", "x", "edge_attr", "scaled_atomic_energies",
            "scaled_total_energy", "total_energy", "forces", "stresses"]


@pytest.mark.parametrize("name", ["wide_lr", "max_lr"])
def test_reference_basis_range(device, name):
    """l_max = 5 / n_max = 6 and l_max = 9 / n_max = 10 (nn/interaction.py:250-253) against the live-reference fixture."""
    from torch_m3gnet_b200 import build_model

    g = golden(name)
    L, R, dim = int(g["l_max"]), int(g["n_max"]), int(g["dim"])
    model = build_model(5.0, 4.0, L, R, 95, dim, 1, device=device)
    model.load_state_dict(state_dict_of(g))
    model.model[6].nsb.factors = torch.from_numpy(g["factors"]).to(device)
    out = model(to_batch(graph_dict(g), device))
    n = g["g.pos"].shape[0]
    report(f"{name}.energy", out["total_energy"], g["out.total_energy"], n * 1e-5, 1e-5)
    report(f"{name}.forces", out["forces"], g["out.forces"], 1e-4, 1e-3)
    for k in ("edge_weights", "x", "edge_attr", "scaled_atomic_energies", "stresses"):
        report(f"{name}.{k}", out[k], g["out." + k], 2e-5, 2e-5)


def test_too_large_basis_raises_like_the_reference(device):
    from torch_m3gnet_b200 import build_model

    with pytest.raises(ValueError):
        build_model(5.0, 4.0, 10, 3, 95, 64, 1, device=device)
    with pytest.raises(ValueError):
        build_model(5.0, 4.0, 3, 11, 95, 64, 1, device=device)


def test_gated_mlp_forward(device):
    """nn/core.py:61-62 called on its own, forward and input gradient, three layer stacks."""
    from torch_m3gnet_b200.nn.core import GatedMLP

    g = golden("gated_mlp")
    for i in range(3):
        cfg = g[f"m{i}.cfg"].tolist()
        fin, is_out, bias, dims = cfg[0], bool(cfg[1]), bool(cfg[2]), cfg[3:]
        mlp = GatedMLP(fin, dims, is_output=is_out, use_bias=bias, device=device)
        mlp.load_state_dict({k[len(f"m{i}.sd."):]: torch.from_numpy(np.array(g[k])) for k in g.files
                             if k.startswith(f"m{i}.sd.")})
        x = torch.from_numpy(g[f"m{i}.x"]).to(device).requires_grad_(True)
        y = mlp(x)
        (gx,) = torch.autograd.grad(y, x, grad_outputs=torch.from_numpy(g[f"m{i}.go"]).to(device))
        report(f"gmlp{i}.y", y, g[f"m{i}.y"], 1e-6, 2e-6)
        report(f"gmlp{i}.gx", gx, g[f"m{i}.gx"], 1e-6, 5e-6)


def test_bonds_not_grouped_by_source_are_regrouped(device):
    """Randomly permuted bonds (live-reference fixture): same results, bond-level outputs in the caller's order."""
    from torch_m3gnet_b200 import build_model
    from torch_m3gnet_b200.data.material_graph import EdgesNotGrouped, get_plan

    g = golden("unsorted_edges")
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    model.load_state_dict(state_dict_of(g))
    for i in (6, 8, 10):
        model.model[i].nsb.factors = torch.from_numpy(g["factors"]).to(device)
    b = to_batch(graph_dict(g), device)
    with pytest.raises(EdgesNotGrouped):
        get_plan(b)
    out = model(b)
    assert out is b and not b["pos"].requires_grad
    report("unsorted.energy", out["total_energy"], g["out.total_energy"], 32 * 1e-5, 1e-5)
    report("unsorted.forces", out["forces"], g["out.forces"], 1e-4, 1e-3)
    for k in ("edge_distances", "triplet_angles", "edge_weights", "edge_attr", "x"):
        report(f"unsorted.{k}", out[k], g["out." + k], 2e-5, 2e-5)
    out2 = model(b)  # twice on the same object (tests/test_model.py:23-34)
    assert torch.equal(out2["forces"], out["forces"])


def test_threebody_interaction_from_public_keys(device):
    """ThreeBodyInteration called on a graph that only carries the public keys (reference nn/interaction.py:187-192):
    the bond vectors are derived from pos / lattice when DistanceAndAngle has not run in this process."""
    from torch_m3gnet_b200.nn.interaction import ThreeBodyInteration

    g = golden("threebody_op")
    gd = graph_dict(g)
    b = to_batch(gd, device)
    tb = ThreeBodyInteration(5.0, 4.0, 3, 3, 64, 64, device=device)
    tb.load_state_dict(state_dict_of(g))
    tb.nsb.factors = torch.from_numpy(g["factors"]).to(device)
    b["x"], b["edge_attr"] = torch.from_numpy(g["x"]).to(device), torch.from_numpy(g["e"]).to(device)
    assert not b._private
    out = tb(b)
    report("tb.public.out", out["edge_attr"], g["out"], 2e-6, 2e-6)
    report("tb.public.dist", out["edge_distances"], g["dist"], 1e-6, 1e-6)
    report("tb.public.cos", out["triplet_angles"], g["cos"], 1e-6, 0)


def test_out_of_range_atom_types_raise(device):
    """The reference's one_hot raises for Z - 1 >= num_types (nn/featurizer.py:36); so do we (no silent table overrun)."""
    from torch_m3gnet_b200 import Batch, build_model

    lat, cart, z = O.fcc_supercell(2, jitter=0.05, seed=0)  # Cu: type 28
    model = build_model(5.0, 4.0, 3, 3, 10, 64, 1, device=device)
    with pytest.raises(ValueError, match="atom_types"):
        model(Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 4.0, device=device))
    gd = graph_dict(golden("c1_default"))
    gd["atom_types"] = gd["atom_types"] - 40  # negative types in a hand-built graph
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 1, device=device)
    with pytest.raises(ValueError, match="atom_types"):
        model(to_batch(gd, device))
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 1, elemental_energies=torch.zeros(5), device=device)
    with pytest.raises(ValueError, match="elemental"):
        model(to_batch(graph_dict(golden("c1_default")), device))


def test_double_backward_raises(device):
    """Training through the forces (create_graph=True, nn/gradient.py:33) is not provided: it must fail loudly."""
    from torch_m3gnet_b200.nn._functions import EdgeAdjustFn

    h = torch.rand(50, 3, device=device, requires_grad=True)
    wt = torch.rand(3, 64, device=device)
    e0 = EdgeAdjustFn.apply(h, wt)
    (gh,) = torch.autograd.grad(e0.sum(), h, create_graph=True)
    # once_differentiable: the first-order gradient is cut off from the graph, so differentiating it again raises
    assert not gh.requires_grad
    with pytest.raises(RuntimeError):
        gh.sum().backward()


def test_graphed_step_notices_parameter_edits(device):
    from torch_m3gnet_b200 import build_model
    from torch_m3gnet_b200.graphed import GraphedStep

    g = golden("c1_default")
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    model.load_state_dict(state_dict_of(g))
    step = GraphedStep(model, to_batch(graph_dict(g), device))
    step()
    assert not step.stale()
    with torch.no_grad():
        model.model[12].gated.dense[0].weight.mul_(1.5)
    assert step.stale()
    with pytest.raises(RuntimeError, match="parameters changed"):
        step()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_kernels_follow_the_tensors_device():
    """Batch on cuda:1 while cuda:0 is the current device: same results as on cuda:0."""
    from torch_m3gnet_b200 import build_model

    g = golden("c1_default")
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=torch.device(dev))
        model.load_state_dict(state_dict_of(g))
        out = model(to_batch(graph_dict(g), torch.device(dev)))
        assert out["forces"].device == torch.device(dev)
        outs.append((out["total_energy"].cpu(), out["forces"].cpu()))
    assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
