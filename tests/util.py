"""Shared helpers for the test-suite (golden fixtures, conversions between oracle dict graphs and Batch)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPH_KEYS = ["pos", "atom_types", "num_triplet_i", "edge_index", "edge_cell_shift", "num_triplet_ij",
              "triplet_edge_index", "lattice", "batch"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def graph_dict(g, prefix="g."):
    """dict of CPU tensors (the oracle's graph format) from a golden file."""
    return {k: torch.from_numpy(np.array(g[prefix + k])) for k in GRAPH_KEYS}


def state_dict_of(g, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(prefix)}


def to_batch(gd, device):
    """Our Batch object from an oracle-format dict."""
    from torch_m3gnet_b200.data.material_graph import Batch

    b = Batch(pos=gd["pos"].clone().to(device), atom_types=gd["atom_types"].to(device),
              num_triplet_i=gd["num_triplet_i"].to(device), edge_index=gd["edge_index"].to(device),
              edge_cell_shift=gd["edge_cell_shift"].to(device), num_triplet_ij=gd["num_triplet_ij"].to(device),
              triplet_edge_index=gd["triplet_edge_index"].to(device), lattice=gd["lattice"].to(device))
    b["batch"] = gd["batch"].to(device)
    return b


def clone_graph(gd):
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in gd.items()}


def report(name, got, want, atol, rtol):
    got = got.detach().cpu().double()
    want = want.detach().cpu().double() if torch.is_tensor(want) else torch.from_numpy(np.array(want)).double()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    diff = (got - want).abs()
    scale = want.abs().max().item() if want.numel() else 0.0
    md = diff.max().item() if diff.numel() else 0.0
    print(f"[parity] {name}: max|ref|={scale:.4e} max|diff|={md:.4e} rel={md / max(scale, 1e-30):.3e}")
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    assert md <= atol + rtol * scale, f"{name}: max diff {md:.3e} > {atol:.1e} + {rtol:.1e}*{scale:.3e}"
