"""TEST INFRASTRUCTURE ONLY — imports the *live* reference package from /root/reference.

The reference (lan496/torch-m3gnet) is pure Python but depends on wheels that are not
installed in this image (torch_scatter, torch_geometric, torchtyping, pymatgen).  This module
installs four tiny stand-in modules in ``sys.modules`` (SURVEY.md §8(c) "shim recipe") so that
``torch_m3gnet.nn.*``, ``torch_m3gnet.model.build`` and
``torch_m3gnet.data.material_graph.compute_threebody`` import and run **unchanged**.

It exists for two purposes only:
  * ``oracle/make_golden.py`` uses it to generate the fixtures under ``tests/golden/``;
  * ``tests/test_oracle_pinned.py`` uses it to pin ``oracle/m3gnet_oracle.py`` against the real thing;
  * ``bench.py --impl reference`` / ``cpu_baseline`` time it on the GPU box's host cores, from the copy that
    ``oracle/build_ref.py`` places under ``oracle/_ref/`` (git-ignored, shipped by gpurun).

Nothing under ``torch_m3gnet_b200/`` may import this file.
"""
from __future__ import annotations

import os
import sys
import types

import torch

# the reference where it lies in the build container, else the copy oracle/build_ref.py ships to the GPU box
_CANDIDATES = ("/root/reference/src", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))


def _find_src():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "torch_m3gnet")):
            return c
    return None


REFERENCE_SRC = _find_src() or _CANDIDATES[0]


def available() -> bool:
    return _find_src() is not None


def location() -> str:
    return _find_src() or "absent"


def _scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    """torch_scatter.scatter_sum semantics (sum ``src`` rows into ``index`` slots along ``dim``)."""
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    index = index.expand_as(src)
    if out is None:
        size = list(src.size())
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        size[dim] = dim_size
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


class _GraphDict(dict):
    """Stand-in for torch_geometric.data.Data: keyword init, item access."""

    def __init__(self, **kw):
        super().__init__(**kw)


def install_shims() -> None:
    if "torch_m3gnet" in sys.modules:
        return
    tt = types.ModuleType("torchtyping")

    class TensorType:  # noqa: D401 - annotation stub
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = TensorType
    sys.modules.setdefault("torchtyping", tt)

    ts = types.ModuleType("torch_scatter")
    ts.scatter_sum = _scatter_sum
    sys.modules.setdefault("torch_scatter", ts)

    pm = types.ModuleType("pymatgen")
    pmc = types.ModuleType("pymatgen.core")

    class Structure:  # noqa: D401 - annotation stub
        pass

    pmc.Structure = Structure
    pm.core = pmc
    sys.modules.setdefault("pymatgen", pm)
    sys.modules.setdefault("pymatgen.core", pmc)

    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.data")
    tgd.Data = _GraphDict
    tgd.InMemoryDataset = object
    tg.data = tgd
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.data", tgd)

    src = _find_src()
    if src is not None and src not in sys.path:
        sys.path.insert(0, src)


def import_reference():
    """Return (build_model, compute_threebody, interaction module, nn package) of the live reference."""
    if not available():
        raise RuntimeError("neither /root/reference nor oracle/_ref is present — live reference unavailable "
                           "(run `python -m oracle.build_ref` in the build container)")
    install_shims()
    from torch_m3gnet.data.material_graph import compute_threebody  # type: ignore
    from torch_m3gnet.model.build import build_model  # type: ignore
    import torch_m3gnet.nn.interaction as interaction  # type: ignore

    return build_model, compute_threebody, interaction


def as_reference_graph(g: dict) -> dict:
    """Plain dict with the keys the reference ``Gradient`` model reads."""
    keys = [
        "pos", "atom_types", "edge_index", "edge_cell_shift", "triplet_edge_index",
        "num_triplet_i", "num_triplet_ij", "lattice", "batch",
    ]
    out = {}
    for k in keys:
        v = g[k]
        out[k] = v.clone() if torch.is_tensor(v) else v
    return out
