"""TEST INFRASTRUCTURE ONLY — round-2 fixtures written from the LIVE reference (same rules as oracle/make_golden.py;
kept in a separate script so that the round-1 fixtures keep regenerating bit for bit).

Run in the build container (``python -m oracle.make_golden_r2``); needs /root/reference.

Fixtures
  wide_lr.npz         l_max = 5, n_max = 6 (outside the round-1 kernel range), dim 64, 1 block, O(1) factor table, on a
                      3-species MPF-like cell: the reference model's outputs
  max_lr.npz          the reference's largest basis l_max = 9, n_max = 10 (nn/interaction.py:250-253), dim 32, 1 block
  gated_mlp.npz       GatedMLP.forward (nn/core.py:61-62) called on its own: three shapes, outputs and input gradients
  unsorted_edges.npz  the default model on a graph whose bonds are NOT grouped by source atom (randomly permuted
                      edge_index / edge_cell_shift / num_triplet_ij, triplet list renumbered): outputs in that order
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import live_reference as lr
from oracle import m3gnet_oracle as O
from oracle.make_golden import OUT, _np, _pack_graph, _pack_sd, _run


def main():
    build_model, compute_threebody, inter = lr.import_reference()
    from torch_m3gnet.nn.core import GatedMLP  # type: ignore
    torch.set_num_threads(1)

    # ---------------- wide_lr / max_lr ----------------
    lat, cart, z = O.mpf_like_structure(2)
    g = O.collate([O.build_graph(lat, cart, z, 5.0, 4.0)])
    for name, (L, R, dim, seed) in {"wide_lr": (5, 6, 64, 21), "max_lr": (9, 10, 32, 22)}.items():
        torch.manual_seed(seed)
        model = build_model(5.0, 4.0, L, R, 95, dim, 1)
        sd = {k: (v.detach() * 2 if k.endswith("weight") else v.detach().clone()) for k, v in model.state_dict().items()}
        model.load_state_dict(sd)
        fac = torch.rand(L, R, generator=torch.Generator().manual_seed(seed)) + 0.5
        model.model[6].nsb.factors = fac
        data = {"l_max": np.array(L), "n_max": np.array(R), "dim": np.array(dim), "factors": _np(fac)}
        data.update(_pack_graph("g.", g))
        data.update(_pack_sd("sd.", sd))
        data.update({f"out.{k}": v for k, v in _run(model, g).items()})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)

    # ---------------- gated_mlp ----------------
    data = {}
    for i, (fin, dims, is_out, bias) in enumerate([(192, [64, 64], False, True), (9, [64], False, False),
                                                   (64, [64, 64, 1], True, True)]):
        torch.manual_seed(30 + i)
        mlp = GatedMLP(fin, dims, is_output=is_out, use_bias=bias)
        x = torch.randn(37, fin, requires_grad=True)
        y = mlp(x)
        go = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, grad_outputs=go)
        data.update(_pack_sd(f"m{i}.sd.", mlp.state_dict()))
        data.update({f"m{i}.x": _np(x), f"m{i}.y": _np(y), f"m{i}.go": _np(go), f"m{i}.gx": _np(gx),
                     f"m{i}.cfg": np.array([fin, int(is_out), int(bias)] + dims)})
    np.savez_compressed(os.path.join(OUT, "gated_mlp.npz"), **data)

    # ---------------- unsorted_edges ----------------
    lat, cart, z = O.fcc_supercell(2, jitter=0.05, seed=6)
    g0 = O.collate([O.build_graph(lat, cart, z, 5.0, 4.0)])
    E = g0["edge_index"].shape[1]
    perm = torch.randperm(E, generator=torch.Generator().manual_seed(6))   # new row r holds old bond perm[r]
    rank = torch.empty(E, dtype=torch.long)
    rank[perm] = torch.arange(E)
    gu = dict(g0)
    gu["edge_index"] = g0["edge_index"][:, perm].contiguous()
    gu["edge_cell_shift"] = g0["edge_cell_shift"][perm].contiguous()
    gu["num_triplet_ij"] = g0["num_triplet_ij"][perm].contiguous()
    gu["triplet_edge_index"] = rank[g0["triplet_edge_index"]].contiguous()
    assert not bool((gu["edge_index"][0][1:] >= gu["edge_index"][0][:-1]).all())
    torch.manual_seed(0)
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3)
    sd3 = {k: (v.detach() * 3 if k.endswith("weight") else v.detach().clone()) for k, v in model.state_dict().items()}
    model.load_state_dict(sd3)
    fac = torch.rand(3, 3, generator=torch.Generator().manual_seed(6)) + 0.5
    for i in (6, 8, 10):
        model.model[i].nsb.factors = fac
    data = {"factors": _np(fac)}
    data.update(_pack_graph("g.", gu))
    data.update(_pack_sd("sd.", sd3))
    data.update({f"out.{k}": v for k, v in _run(model, gu).items()})
    # the same model on the grouped graph: bond-level outputs must be the permuted rows, the rest identical
    ref = _run(model, g0)
    assert np.array_equal(ref["edge_distances"][perm.numpy()], data["out.edge_distances"])
    assert np.allclose(ref["forces"], data["out.forces"], rtol=0, atol=2e-6)
    np.savez_compressed(os.path.join(OUT, "unsorted_edges.npz"), **data)
    for f in ("wide_lr.npz", "max_lr.npz", "gated_mlp.npz", "unsorted_edges.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
