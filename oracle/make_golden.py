"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz from the LIVE reference.

Run in the build container (``python -m oracle.make_golden``); needs /root/reference.  Every output
array in the fixtures is produced by the reference's own, unmodified modules (imported through the
shims of ``oracle/live_reference.py``) running on CPU in fp32.  The *inputs* (graphs) come from
``oracle.m3gnet_oracle.build_graph`` because the reference's neighbour search lives in pymatgen,
which is not installed; the triplet lists inside those graphs are cross-checked here against the
reference's own ``compute_threebody``.

Fixtures
  c1_default.npz      config 1 of BASELINE.json: 2x2x2 FCC Cu (+-0.05 A, default_rng(0)), default model,
                      torch.manual_seed(0) weights; also the same with all *.weight x 3
  small_batch.npz     the reference test-suite's model (l_max=2,n_max=3,num_types=93,dim=17,blocks=2;
                      tests/conftest.py:150-178) on its FCC-Al + BCC-Na batch, positions perturbed
  tio2_default.npz    the 32-atom Ti8O24 cell of tests/conftest.py:45-86 (two species), default model
  threebody_op.npz    operator-level ThreeBodyInteration fwd/bwd with an injected O(1) factor table and
                      random non-unit upstream gradients (quirks Q1/Q3)
  conv_op.npz         operator-level M3GNetConv fwd/bwd
  basis.npz           known answers of the basis functions incl. the Legendre backward quirk
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import live_reference as lr
from oracle import m3gnet_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

GRAPH_KEYS = ["pos", "atom_types", "num_triplet_i", "edge_index", "edge_cell_shift", "num_triplet_ij",
              "triplet_edge_index", "lattice", "batch"]
OUT_KEYS = ["edge_distances", "triplet_angles", "edge_weights", "x", "edge_attr",
            "scaled_atomic_energies", "scaled_total_energy", "total_energy", "forces", "stresses"]


def _np(t):
    return t.detach().cpu().numpy()


def _pack_graph(prefix, g):
    return {f"{prefix}{k}": _np(g[k]) for k in GRAPH_KEYS}


def _pack_sd(prefix, sd):
    return {f"{prefix}{k}": _np(v) for k, v in sd.items()}


def _run(model, g):
    out = model(lr.as_reference_graph(g))
    return {k: _np(out[k]) for k in OUT_KEYS}


def _check_triplets(compute_threebody, g, n, r3):
    tri, nti, ntij = compute_threebody(n, g["edge_index"], g["edge_distances_build"], r3)
    assert torch.equal(tri, g["triplet_edge_index"])
    assert torch.equal(nti, g["num_triplet_i"]) and torch.equal(ntij, g["num_triplet_ij"])


def main():
    os.makedirs(OUT, exist_ok=True)
    build_model, compute_threebody, inter = lr.import_reference()
    torch.set_num_threads(1)  # deterministic CPU reductions

    # ---------------- c1_default ----------------
    lat, cart, z = O.fcc_supercell(2, jitter=0.05, seed=0)
    g = O.build_graph(lat, cart, z, 5.0, 4.0)
    _check_triplets(compute_threebody, g, len(cart), 4.0)
    b = O.collate([g])
    torch.manual_seed(0)
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    data = {}
    data.update(_pack_graph("g.", b))
    data.update(_pack_sd("sd.", sd))
    data.update({f"out.{k}": v for k, v in _run(model, b).items()})
    data["factors"] = _np(model.model[6].nsb.factors)
    # amplified weights (x3 on every *.weight) — forces become O(0.1) eV/A
    sd3 = {k: (v * 3 if k.endswith("weight") else v.clone()) for k, v in sd.items()}
    model.load_state_dict(sd3)
    data.update({f"out3.{k}": v for k, v in _run(model, b).items()})
    np.savez_compressed(os.path.join(OUT, "c1_default.npz"), **data)

    # ---------------- tio2_default (two species, reference fixture geometry) ----------------
    a = 8.01
    coords = np.array([
        [0.005698, 7.903250, 7.975364], [7.962333, 0.031776, 4.087014], [7.987993, 4.053572, 7.916418],
        [7.972553, 3.990096, 3.904352], [3.901632, 0.009469, 0.015298], [4.061435, 7.980741, 3.923483],
        [4.075226, 3.974756, 0.060859], [3.997434, 3.997462, 3.900065], [0.002131, 2.089909, 2.043724],
        [7.935880, 2.054631, 6.053889], [7.986174, 5.996277, 1.901030], [0.073084, 5.950515, 5.952990],
        [4.057353, 2.078078, 1.975213], [4.049787, 2.018112, 6.084813], [3.971569, 5.919147, 2.051521],
        [3.945378, 6.072591, 6.041797], [1.964716, 0.069527, 2.062618], [1.928378, 7.984901, 6.068134],
        [1.990663, 4.042357, 2.090104], [1.974315, 3.921490, 6.056360], [6.008068, 7.938413, 2.078371],
        [5.953855, 0.062646, 6.062819], [5.900438, 4.009349, 1.999860], [6.040758, 3.924354, 6.051151],
        [1.936480, 1.932966, 0.038363], [2.043398, 1.921099, 3.956512], [1.983471, 5.951049, 0.085619],
        [2.010997, 6.095910, 4.026083], [5.955844, 1.984438, 7.911637], [6.075395, 1.996245, 4.065586],
        [6.080717, 5.987091, 7.942396], [5.983861, 5.933218, 3.927338]])
    zz = np.array([22] * 8 + [8] * 24)
    g2 = O.build_graph(np.eye(3) * a, coords, zz, 5.0, 4.0)
    _check_triplets(compute_threebody, g2, 32, 4.0)
    b2 = O.collate([g2])
    model.load_state_dict(sd3)  # amplified default weights: larger, species-dependent forces
    data = {}
    data.update(_pack_graph("g.", b2))
    data.update({f"out3.{k}": v for k, v in _run(model, b2).items()})
    model.load_state_dict(sd)
    data.update({f"out.{k}": v for k, v in _run(model, b2).items()})
    np.savez_compressed(os.path.join(OUT, "tio2_default.npz"), **data)

    # ---------------- small_batch (reference test model on FCC-Al + BCC-Na) ----------------
    r_nn = 3.0
    lat_al = r_nn * np.sqrt(2) * np.eye(3)
    fr_al = np.array([[0, 0, 0], [0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]])
    lat_na = r_nn / np.sqrt(3) * 2 * np.eye(3)
    fr_na = np.array([[0, 0, 0], [0.5, 0.5, 0.5]])
    rc = r_nn + 1e-4
    ga = O.build_graph(lat_al, fr_al @ lat_al, [13] * 4, rc, rc)
    gn = O.build_graph(lat_na, fr_na @ lat_na, [11] * 2, rc, rc)
    assert ga["num_triplet_i"].tolist() == [132] * 4 and gn["num_triplet_i"].tolist() == [56] * 2
    rng = np.random.default_rng(7)
    for gg in (ga, gn):  # perturb after the graph is built, as tests/test_model.py:62-66 does
        gg["pos"] = gg["pos"] + torch.tensor(0.1 * (rng.random(gg["pos"].shape) - 0.5), dtype=torch.float)
    bs = O.collate([ga, gn])
    torch.manual_seed(1)
    small = build_model(rc, rc, 2, 3, 93, 17, 2)
    sds = {k: v.detach().clone() for k, v in small.state_dict().items()}
    data = {}
    data.update(_pack_graph("g.", bs))
    data.update(_pack_sd("sd.", sds))
    data.update({f"out.{k}": v for k, v in _run(small, bs).items()})
    data["cutoff"] = np.array(rc)
    data["factors"] = _np(small.model[6].nsb.factors)
    np.savez_compressed(os.path.join(OUT, "small_batch.npz"), **data)

    # ---------------- threebody_op ----------------
    from torch_m3gnet.nn.interaction import ThreeBodyInteration  # type: ignore
    from torch_m3gnet.nn.conv import M3GNetConv  # type: ignore

    lat, cart, z = O.fcc_supercell(2, jitter=0.1, seed=3)
    g3 = O.collate([O.build_graph(lat, cart, z, 5.0, 4.0)])
    torch.manual_seed(3)
    tb = ThreeBodyInteration(5.0, 4.0, 3, 3, 64, 64)
    fac = torch.rand(3, 3) + 0.5
    tb.nsb.factors = fac
    N, E = g3["pos"].shape[0], g3["edge_index"].shape[1]
    x = (0.5 * torch.randn(N, 64)).requires_grad_(True)
    e = (0.5 * torch.randn(E, 64)).requires_grad_(True)
    vec, dist, cos = O.pair_geometry(g3["pos"], g3["lattice"], g3["batch"], g3["edge_index"],
                                     g3["edge_cell_shift"], g3["triplet_edge_index"])
    dist = dist.detach().requires_grad_(True)
    cos = cos.detach().requires_grad_(True)
    graph = {"edge_distances": dist, "triplet_angles": cos, "x": x, "edge_attr": e,
             "edge_index": g3["edge_index"], "triplet_edge_index": g3["triplet_edge_index"]}
    captured = {}
    hook = tb.gated_mlp.register_forward_hook(lambda m, i, o: captured.__setitem__("red", i[0].detach().clone()))
    out = tb(dict(graph))["edge_attr"]
    hook.remove()
    go = torch.randn(E, 64)
    gx, ge, gr, gc = torch.autograd.grad(out, [x, e, dist, cos], grad_outputs=go)
    data = {}
    data.update(_pack_graph("g.", g3))
    data.update(_pack_sd("sd.", tb.state_dict()))
    data.update(factors=_np(fac), x=_np(x), e=_np(e), dist=_np(dist), cos=_np(cos), vec=_np(vec),
                red=_np(captured["red"]), out=_np(out), go=_np(go), gx=_np(gx), ge=_np(ge), gr=_np(gr), gc=_np(gc))
    np.savez_compressed(os.path.join(OUT, "threebody_op.npz"), **data)

    # ---------------- conv_op ----------------
    torch.manual_seed(4)
    cv = M3GNetConv(3, 64, 64)
    x = (0.5 * torch.randn(N, 64)).requires_grad_(True)
    e = (0.5 * torch.randn(E, 64)).requires_grad_(True)
    h = (0.3 * torch.randn(E, 3)).requires_grad_(True)
    res = cv({"x": x, "edge_attr": e, "edge_weights": h, "edge_index": g3["edge_index"]})
    gox, goe = torch.randn(N, 64), torch.randn(E, 64)
    gx, ge, gh = torch.autograd.grad([res["x"], res["edge_attr"]], [x, e, h], grad_outputs=[gox, goe])
    data = {"edge_index": _np(g3["edge_index"])}
    data.update(_pack_sd("sd.", cv.state_dict()))
    data.update(x=_np(x), e=_np(e), h=_np(h), x_out=_np(res["x"]), e_out=_np(res["edge_attr"]),
                gox=_np(gox), goe=_np(goe), gx=_np(gx), ge=_np(ge), gh=_np(gh))
    np.savez_compressed(os.path.join(OUT, "conv_op.npz"), **data)

    # ---------------- basis ----------------
    r = torch.tensor([1.0, 2.556, 3.615, 4.0, 4.427, 5.0])
    from torch_m3gnet.nn.featurizer import EdgeFeaturizer  # type: ignore
    ef = EdgeFeaturizer(3, 5.0)
    h = ef({"edge_distances": r})["edge_weights"]
    xs = torch.tensor([0.3, -0.7, 1.0, -1.0, 0.0], requires_grad=True)
    leg = {}
    for l in range(4):
        y = inter.legendre_cos(xs, l)
        (gl,) = torch.autograd.grad(y, xs, grad_outputs=torch.full_like(xs, 0.5))
        leg[f"leg{l}"] = _np(y)
        leg[f"leg{l}_grad_go0.5"] = _np(gl)
    xb = torch.linspace(0.1, 14.0, 40, requires_grad=True)
    bes = {}
    for l in range(4):
        y = inter.spherical_bessel(xb, l)
        (gl,) = torch.autograd.grad(y, xb, grad_outputs=torch.ones_like(xb))
        bes[f"j{l}"] = _np(y)
        bes[f"j{l}_grad"] = _np(gl)
    np.savez_compressed(
        os.path.join(OUT, "basis.npz"), r=_np(r), h=_np(h), em=_np(ef.em), dm=_np(ef.dm), coeff=_np(ef.coeff),
        fc4=_np(inter.cutoff_function(torch.tensor([1.0, 2.556, 3.615, 4.0, 4.5]), 4.0)),
        leg_x=_np(xs), bes_x=_np(xb), zeros=np.array(inter.SPHERICAL_BESSEL_ZEROS), **leg, **bes)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
