"""CPU: the oracle restatement against the fixtures written by the LIVE reference (oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import m3gnet_oracle as O
from tests.util import clone_graph, golden, graph_dict, report, state_dict_of

OUT_KEYS = ["edge_distances", "triplet_angles", "edge_weights", "x", "edge_attr", "scaled_atomic_energies",
            "scaled_total_energy", "total_energy", "forces", "stresses"]


def _check(out, g, prefix, rtol=2e-6, atol=1e-9):
    for k in OUT_KEYS:
        report(prefix + k, out[k], g[prefix + k], atol=atol, rtol=rtol)


def test_c1_default_and_amplified():
    torch.set_num_threads(1)
    g = golden("c1_default")
    gd, sd = graph_dict(g), state_dict_of(g)
    hp = O.HyperParams()
    fac = torch.from_numpy(g["factors"])
    _check(O.forward(sd, hp, clone_graph(gd), factors=fac), g, "out.")
    sd3 = {k: (v * 3 if k.endswith("weight") else v) for k, v in sd.items()}
    _check(O.forward(sd3, hp, clone_graph(gd), factors=fac), g, "out3.")
    assert abs(float(g["out.total_energy"][0]) - (-1.0142778)) < 1e-6  # SURVEY §8(c) known answer


def test_tio2_two_species():
    torch.set_num_threads(1)
    g, c1 = golden("tio2_default"), golden("c1_default")
    gd, sd = graph_dict(g), state_dict_of(c1)
    hp = O.HyperParams()
    fac = torch.from_numpy(c1["factors"])
    _check(O.forward(sd, hp, clone_graph(gd), factors=fac), g, "out.")
    sd3 = {k: (v * 3 if k.endswith("weight") else v) for k, v in sd.items()}
    _check(O.forward(sd3, hp, clone_graph(gd), factors=fac), g, "out3.")


def test_small_batch_reference_test_model():
    torch.set_num_threads(1)
    g = golden("small_batch")
    rc = float(g["cutoff"])
    hp = O.HyperParams(cutoff=rc, threebody_cutoff=rc, l_max=2, n_max=3, num_types=93, embedding_dim=17, num_blocks=2)
    out = O.forward(state_dict_of(g), hp, clone_graph(graph_dict(g)), factors=torch.from_numpy(g["factors"]))
    _check(out, g, "out.")
    # reference tests/test_data.py:18-23 known answer: 12*11 (FCC) and 8*7 (BCC) triplets per atom
    assert g["g.num_triplet_i"].tolist() == [132, 132, 132, 132, 56, 56]
    assert g["g.batch"].tolist() == [0, 0, 0, 0, 1, 1]


def test_threebody_operator_with_injected_factors():
    torch.set_num_threads(1)
    g = golden("threebody_op")
    sd = {"tb." + k: v for k, v in state_dict_of(g).items()}
    hp = O.HyperParams()
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    e = torch.from_numpy(g["e"]).requires_grad_(True)
    dist = torch.from_numpy(g["dist"]).requires_grad_(True)
    cos = torch.from_numpy(g["cos"]).requires_grad_(True)
    gd = graph_dict(g)
    out, red = O.three_body(sd, "tb", hp, x, e, dist, cos, gd["edge_index"], gd["triplet_edge_index"],
                            torch.from_numpy(g["factors"]))
    report("tb.red", red, g["red"], atol=1e-9, rtol=2e-6)
    report("tb.out", out, g["out"], atol=1e-9, rtol=2e-6)
    gx, ge, gr, gc = torch.autograd.grad(out, [x, e, dist, cos], grad_outputs=torch.from_numpy(g["go"]))
    for name, got in (("gx", gx), ("ge", ge), ("gr", gr), ("gc", gc)):
        report("tb." + name, got, g[name], atol=1e-8, rtol=5e-6)
    # 810 of the 4 < r <= 5 edges carry no triplets: their reduced features are exactly zero
    assert (np.abs(g["red"]).sum(axis=1) == 0).sum() > 0


def test_conv_operator():
    torch.set_num_threads(1)
    g = golden("conv_op")
    sd = {"cv." + k: v for k, v in state_dict_of(g).items()}
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    e = torch.from_numpy(g["e"]).requires_grad_(True)
    h = torch.from_numpy(g["h"]).requires_grad_(True)
    x2, e2 = O.conv(sd, "cv", x, e, h, torch.from_numpy(g["edge_index"]))
    report("conv.x_out", x2, g["x_out"], atol=1e-9, rtol=2e-6)
    report("conv.e_out", e2, g["e_out"], atol=1e-9, rtol=2e-6)
    gx, ge, gh = torch.autograd.grad([x2, e2], [x, e, h],
                                     grad_outputs=[torch.from_numpy(g["gox"]), torch.from_numpy(g["goe"])])
    for name, got in (("gx", gx), ("ge", ge), ("gh", gh)):
        report("conv." + name, got, g[name], atol=1e-8, rtol=5e-6)


def test_basis_known_answers():
    g = golden("basis")
    r = torch.from_numpy(g["r"])
    report("radial.h", O.radial_basis(r, 3, 5.0), g["h"], atol=1e-9, rtol=1e-6)
    em, dm, coeff = O.radial_constants(3, 5.0)
    assert np.array_equal(em.numpy(), g["em"]) and np.array_equal(dm.numpy(), g["dm"])
    assert np.array_equal(coeff.numpy(), g["coeff"])
    report("fc", O.cutoff_function(torch.tensor([1.0, 2.556, 3.615, 4.0, 4.5]), 4.0), g["fc4"], 1e-9, 1e-6)
    xs = torch.from_numpy(g["leg_x"]).requires_grad_(True)
    xb = torch.from_numpy(g["bes_x"]).requires_grad_(True)
    for l in range(4):
        y = O.legendre_cos(xs, l)
        (gl,) = torch.autograd.grad(y, xs, grad_outputs=torch.full_like(xs, 0.5))
        report(f"leg{l}", y, g[f"leg{l}"], 1e-9, 1e-6)
        report(f"leg{l}.grad", gl, g[f"leg{l}_grad_go0.5"], 1e-9, 1e-6)
        y = O.spherical_bessel(xb, l)
        (gl,) = torch.autograd.grad(y, xb, grad_outputs=torch.ones_like(xb))
        report(f"j{l}", y, g[f"j{l}"], 1e-9, 1e-6)
        report(f"j{l}.grad", gl, g[f"j{l}_grad"], 1e-9, 1e-6)
    # quirk Q3 known answer: legendre_cos([0.3,-0.7], 2) with go=0.5 → [0.375, -0.875]
    assert np.allclose(g["leg2_grad_go0.5"][:2], [0.375, -0.875], atol=1e-6)
    # zeros table: the oracle's scipy derivation equals the reference literals after float32 rounding
    assert np.array_equal(O.bessel_zero_table().astype(np.float32), g["zeros"].astype(np.float32))


def test_neighbor_list_known_answers():
    """Reference tests that pin the pymatgen boundary (tests/test_data.py:18-23, tests/test_nn.py:16-30)."""
    # 1 Å cube, cutoff 1 Å: 4 atoms; self-image edges exist (d == r is included by d^2 < r^2 + 1e-8)
    lat = np.eye(3)
    frac = np.array([[0, 0, 0], [0.5, 0.5, 0.5], [0.5, 0, 0], [0, 0.5, 0.5]])
    src, dst, img, dist = O.neighbor_list_bruteforce(lat, frac @ lat, 1.0)
    assert ((src == dst) & (np.abs(img).sum(axis=1) == 1)).sum() == 6 * 4  # six unit self-images per atom
    assert np.all(dist <= 1.0 + 1e-8) and np.all(dist > 0)
    assert np.all(np.diff(src) >= 0)
    # symmetric list: (i,j,s) present <=> (j,i,-s) present
    fwd = set(zip(src.tolist(), dst.tolist(), map(tuple, img.tolist())))
    assert all((j, i, tuple(-np.array(s))) in fwd for (i, j, s) in fwd)


def test_rotation_invariance_of_distance_multiset():
    """tests/test_invariance.py:41-66 restated: sorted distances of a sheared cell are rotation invariant."""
    g = golden("tio2_default")
    pos = g["g.pos"].astype(np.float64)
    lat = g["g.lattice"][0].astype(np.float64)
    strain = np.eye(3) + 0.1 * np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1.0]])
    frac = pos @ np.linalg.inv(lat)
    lat2 = lat @ strain
    pos2 = frac @ lat2
    c, s = 0.5, np.sqrt(3) / 2
    rot = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]]) @ np.array(
        [[0, 0, 1], [1 / np.sqrt(2), -1 / np.sqrt(2), 0], [1 / np.sqrt(2), 1 / np.sqrt(2), 0]])
    d1 = np.sort(O.neighbor_list_bruteforce(lat2, pos2, 5.0)[3])
    d2 = np.sort(O.neighbor_list_bruteforce(lat2 @ rot.T, pos2 @ rot.T, 5.0)[3])
    assert d1.shape == d2.shape and np.allclose(d1, d2, atol=1e-9)


def _model_fixture(name):
    g = golden(name)
    hp = O.HyperParams(l_max=int(g["l_max"]), n_max=int(g["n_max"]), embedding_dim=int(g["dim"]), num_blocks=1)
    return g, hp


def test_wide_and_maximal_basis_sizes():
    """l_max = 5 / n_max = 6 and the reference's largest basis l_max = 9 / n_max = 10 (nn/interaction.py:250-253)."""
    torch.set_num_threads(1)
    for name in ("wide_lr", "max_lr"):
        g, hp = _model_fixture(name)
        out = O.forward(state_dict_of(g), hp, clone_graph(graph_dict(g)), factors=torch.from_numpy(g["factors"]))
        for k in OUT_KEYS:
            report(f"{name}.{k}", out[k], g["out." + k], atol=2e-8, rtol=5e-6)


def test_gated_mlp_called_on_its_own():
    """GatedMLP.forward (nn/core.py:61-62): three layer stacks incl. the is_output / bias-free variants."""
    g = golden("gated_mlp")
    for i in range(3):
        cfg = g[f"m{i}.cfg"].tolist()
        fin, is_out, bias, dims = cfg[0], bool(cfg[1]), bool(cfg[2]), cfg[3:]
        sd = {"m." + k[len(f"m{i}.sd."):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(f"m{i}.sd.")}
        x = torch.from_numpy(g[f"m{i}.x"]).requires_grad_(True)
        y = O.gated_mlp(sd, "m", x, len(dims), is_output=is_out, bias=bias)
        (gx,) = torch.autograd.grad(y, x, grad_outputs=torch.from_numpy(g[f"m{i}.go"]))
        report(f"gmlp{i}.y", y, g[f"m{i}.y"], 1e-8, 2e-6)
        report(f"gmlp{i}.gx", gx, g[f"m{i}.gx"], 1e-8, 5e-6)


def test_bonds_not_grouped_by_source():
    """The reference accepts any bond order (scatter by index); fixture: randomly permuted bonds."""
    torch.set_num_threads(1)
    g = golden("unsorted_edges")
    gd = graph_dict(g)
    assert not bool((gd["edge_index"][0][1:] >= gd["edge_index"][0][:-1]).all())
    out = O.forward(state_dict_of(g), O.HyperParams(), clone_graph(gd), factors=torch.from_numpy(g["factors"]))
    for k in OUT_KEYS:
        report(f"unsorted.{k}", out[k], g["out." + k], atol=2e-8, rtol=5e-6)
