"""Graph / batch data objects of the reference API (torch_m3gnet/data/material_graph.py:14-165), rebuilt
without torch_geometric, plus the GPU graph builder and the per-batch *plan* (the int32 CSR views the
kernels consume).

* ``MaterialGraph`` / ``Batch``: item + attribute access, ``.to()``, ``.clone()``, ``Batch.from_data_list``
  following the collate rules of material_graph.py:109-130.
* ``MaterialGraph.from_structure`` / ``Batch.from_structures``: neighbour list + triplets on the GPU
  (csrc/neighbor.cu) — one call for any number of structures.
* ``GraphPlan``: canonical form cached on the batch object and invalidated when a structural tensor is
  replaced or modified in place (tests/test_model.py:26-34 permutes ``triplet_edge_index`` in place).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from torch_m3gnet_b200 import _lib
from torch_m3gnet_b200.data import MaterialGraphKey as K

_STRUCTURAL = (K.EDGE_INDEX, K.TRIPLET_EDGE_INDEX, K.EDGE_CELL_SHIFT, K.ATOM_TYPES, K.BATCH)
_GRAPH_LEVEL = (K.LATTICE, K.TOTAL_ENERGY, K.STRESSES)  # new leading batch dimension when collated
_COUNTS = (K.NUM_NODES, K.NUM_EDGES, K.NUM_TRIPLETS)
_DERIVED_DEFAULTS = (
    K.SCALED_POS, K.SCALED_LATTICE, K.EDGE_DISTANCES, K.TRIPLET_ANGLES, K.EDGE_WEIGHTS, K.ELEMENTAL_ENERGIES,
    K.NODE_FEATURES, K.EDGE_ATTR, K.SCALED_ATOMIC_ENERGIES, K.SCALED_TOTAL_ENERGY, K.TOTAL_ENERGY, K.FORCES,
    K.STRESSES,
)


class MaterialGraph:
    """One periodic structure as a graph (fields: material_graph.py:15-60)."""

    def __init__(self, pos=None, atom_types=None, num_triplet_i=None, edge_index=None, edge_cell_shift=None,
                 num_triplet_ij=None, triplet_edge_index=None, lattice=None, **extra):
        store: Dict[str, Any] = {}
        store[K.POS] = pos
        store[K.ATOM_TYPES] = atom_types
        store[K.NUM_TRIPLET_I] = num_triplet_i
        store[K.EDGE_INDEX] = edge_index
        store[K.EDGE_CELL_SHIFT] = edge_cell_shift
        store[K.NUM_TRIPLET_IJ] = num_triplet_ij
        store[K.TRIPLET_EDGE_INDEX] = triplet_edge_index
        store[K.LATTICE] = lattice
        store[K.NUM_NODES] = pos.size(0) if pos is not None else 0
        store[K.NUM_EDGES] = edge_index.size(1) if edge_index is not None else 0
        store[K.NUM_TRIPLETS] = triplet_edge_index.size(1) if triplet_edge_index is not None else 0
        for k in _DERIVED_DEFAULTS:
            store[k] = None
        store.update(extra)
        object.__setattr__(self, "_store", store)
        object.__setattr__(self, "_plan", None)
        object.__setattr__(self, "_private", {})  # kernel-side intermediates (not part of the API)

    # ---- mapping / attribute protocol ----
    def __getitem__(self, key: str):
        return self._store[key]

    def __setitem__(self, key: str, value):
        self._store[key] = value
        if key in _STRUCTURAL:
            object.__setattr__(self, "_plan", None)

    def __contains__(self, key):
        return key in self._store

    def __getattr__(self, name):
        store = object.__getattribute__(self, "_store")
        if name in store:
            return store[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def keys(self):
        return [k for k, v in self._store.items() if v is not None]

    def _new_like(self, store):
        out = object.__new__(type(self))
        object.__setattr__(out, "_store", store)
        object.__setattr__(out, "_plan", None)
        object.__setattr__(out, "_private", {})
        # a no-op move (same device) keeps the same index tensors: the cached plan stays valid
        plan = self._plan
        if plan is not None and plan.signature == GraphPlan.signature_of(out):
            object.__setattr__(out, "_plan", plan)
        return out

    def to(self, device, *args, **kwargs):
        store = {k: (v.to(device, *args, **kwargs) if torch.is_tensor(v) else v) for k, v in self._store.items()}
        return self._new_like(store)

    def clone(self):
        store = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._store.items()}
        return self._new_like(store)

    def cuda(self):
        return self.to(torch.device("cuda"))

    def cpu(self):
        return self.to(torch.device("cpu"))

    def __repr__(self):
        parts = []
        for k, v in self._store.items():
            if torch.is_tensor(v):
                parts.append(f"{k}={list(v.shape)}")
            elif v is not None:
                parts.append(f"{k}={v}")
        return f"{type(self).__name__}({', '.join(parts)})"

    # ---- construction ----
    @classmethod
    def from_structure(cls, structure, cutoff: float, threebody_cutoff: float,
                       device: Optional[torch.device] = None) -> "MaterialGraph":
        """material_graph.py:132-165, with the neighbour search and the triplet loop on the GPU.

        Returned tensors live on the CUDA device (the reference returns CPU tensors and the caller moves them;
        ``.to(device)`` on the result is then a no-op)."""
        batch = Batch.from_structures([structure], cutoff, threebody_cutoff, device=device)
        store = dict(batch._store)
        store.pop(K.BATCH, None)
        store[K.LATTICE] = store[K.LATTICE][0]
        g = object.__new__(cls)
        object.__setattr__(g, "_store", store)
        object.__setattr__(g, "_plan", None)
        object.__setattr__(g, "_private", {})
        return g


class Batch(MaterialGraph):
    """Several graphs concatenated (PyG ``Batch`` under the rules of material_graph.py:109-130)."""

    @property
    def num_graphs(self) -> int:
        return int(self._store[K.LATTICE].size(0))

    @classmethod
    def from_data_list(cls, graphs: Sequence[MaterialGraph]) -> "Batch":
        if len(graphs) == 0:
            raise ValueError("empty data list")
        store: Dict[str, Any] = {}
        n_off, e_off = 0, 0
        edge_index, tri_index, batch = [], [], []
        device = graphs[0][K.POS].device
        for b, g in enumerate(graphs):
            n, e = int(g[K.NUM_NODES]), int(g[K.NUM_EDGES])
            edge_index.append(g[K.EDGE_INDEX] + n_off)
            tri_index.append(g[K.TRIPLET_EDGE_INDEX] + e_off)
            batch.append(torch.full((n,), b, dtype=torch.long, device=device))
            n_off += n
            e_off += e
        store[K.EDGE_INDEX] = torch.cat(edge_index, dim=1)
        store[K.TRIPLET_EDGE_INDEX] = torch.cat(tri_index, dim=1)
        store[K.BATCH] = torch.cat(batch)
        keys = list(graphs[0]._store.keys())
        for k in keys:
            if k in (K.EDGE_INDEX, K.TRIPLET_EDGE_INDEX):
                continue
            vals = [g._store.get(k) for g in graphs]
            if k in _COUNTS:
                store[k] = int(sum(vals))
            elif any(v is None for v in vals):
                store[k] = None
            elif k in _GRAPH_LEVEL:
                store[k] = torch.stack([v.reshape(3, 3) if k == K.LATTICE else v for v in vals])
            elif torch.is_tensor(vals[0]):
                store[k] = torch.cat(vals, dim=0)
            else:
                store[k] = vals
        out = object.__new__(cls)
        object.__setattr__(out, "_store", store)
        object.__setattr__(out, "_plan", None)
        object.__setattr__(out, "_private", {})
        return out

    @classmethod
    def from_structures(cls, structures: Sequence, cutoff: float, threebody_cutoff: float,
                        device: Optional[torch.device] = None) -> "Batch":
        """Batched GPU graph build: neighbour list (m3g_nbr_count/fill) + triplets (m3g_triplet_count/fill)."""
        if threebody_cutoff > cutoff:
            raise ValueError("Three body cutoff raidus should be smaller than two body.")
        lat = np.stack([np.asarray(s.lattice.matrix, dtype=np.float64) for s in structures])
        sizes = [len(s) for s in structures]
        cart = np.concatenate([np.asarray(s.cart_coords, dtype=np.float64).reshape(-1, 3) for s in structures])
        z = np.concatenate([np.array([site.specie.Z for site in s], dtype=np.int64) for s in structures])
        return cls.from_arrays(lat, cart, z, sizes, cutoff, threebody_cutoff, device=device)

    @staticmethod
    def _cell_list(lattices, B, N, cutoff, lat64, cart64, atom_ptr, device):
        """Bins for structures whose cell holds >= 3 bins of perpendicular width >= cutoff along every axis (large
        supercells); small cells keep the plain sweep over the structure's atoms."""
        lat = np.ascontiguousarray(lattices, dtype=np.float64).reshape(B, 3, 3)
        inv = np.linalg.inv(lat)  # columns b_k; 1/|b_k| = spacing of the lattice planes along axis k
        width = 1.0 / np.linalg.norm(inv, axis=1)
        nb = np.minimum(np.floor(width / (cutoff * 1.0001 + 1e-6)), 256).astype(np.int64)
        nb[(nb < 3).any(axis=1)] = 0
        if not nb.any():
            return None, None, None, None
        i32 = dict(dtype=torch.int32, device=device)
        base_h = np.concatenate([[0], np.cumsum(nb.prod(axis=1))]).astype(np.int32)
        n_bins = int(base_h[-1])
        bins = torch.as_tensor(nb.astype(np.int32)).to(device)
        bin_base = torch.as_tensor(base_h).to(device)
        atom_bin = torch.empty(N, **i32)
        bin_count = torch.zeros(n_bins, **i32)
        _lib.call("nbr_bin_count", lat64, cart64, atom_ptr, B, N, cutoff, bins, bin_base, atom_bin, bin_count)
        bin_ptr = torch.empty(n_bins + 1, **i32)
        work = torch.empty(_lib.scan_work_elems(n_bins), **i32)
        _lib.call("exclusive_scan_i32", bin_count, bin_ptr, n_bins, work)
        bin_cursor = torch.zeros(n_bins, **i32)
        bin_atoms = torch.empty(N, **i32)
        _lib.call("nbr_bin_fill", atom_bin, bin_ptr, N, bin_cursor, bin_atoms)
        return bins, bin_base, bin_ptr, bin_atoms

    @classmethod
    def _sweep(cls, lattices, lat64, cart64, atom_ptr, B, N, cutoff, threebody_cutoff, device):
        """Neighbour sweep (count -> scan -> fill) of the atoms in ``cart64``; returns the source-CSR offsets, the
        bond count and the bond tensors."""
        i32 = dict(dtype=torch.int32, device=device)
        bins, bin_base, bin_ptr, bin_atoms = cls._cell_list(lattices, B, N, float(cutoff), lat64, cart64, atom_ptr,
                                                            device)
        counts = torch.empty(N, **i32)
        _lib.call("nbr_count", lat64, cart64, atom_ptr, B, N, float(cutoff), bins, bin_base, bin_ptr, bin_atoms,
                  counts)
        edge_ptr = torch.empty(N + 1, **i32)
        work = torch.empty(_lib.scan_work_elems(N), **i32)
        _lib.call("exclusive_scan_i32", counts, edge_ptr, N, work)
        E = int(edge_ptr[-1].item())
        edge_index = torch.empty((2, E), dtype=torch.int64, device=device)
        shift = torch.empty((E, 3), **i32)
        dist = torch.empty(E, dtype=torch.float32, device=device)
        member = torch.empty(E, **i32)
        _lib.call("nbr_fill", lat64, cart64, atom_ptr, B, N, float(cutoff), float(threebody_cutoff), bins, bin_base,
                  bin_ptr, bin_atoms, edge_ptr, E, edge_index, shift, dist, member)
        return edge_ptr, E, edge_index, shift, dist, member

    @classmethod
    def _assemble(cls, lat64, cart64, types, atom_ptr, batch, N, edge_ptr, E, edge_index, shift, dist, member,
                  want_triplet_index, lat32=None, type_range=None):
        """Triplets of the bond list + the Batch object with its plan seeded from the builder's CSR."""
        device = cart64.device
        i32 = dict(dtype=torch.int32, device=device)
        nti = torch.empty(N, dtype=torch.int64, device=device)
        ntij = torch.empty(E, **i32)
        tri_count = torch.empty(E, **i32)
        member_list = torch.empty(max(E, 1), **i32)
        used_count = torch.empty(N, **i32)
        stats = torch.zeros(3, dtype=torch.int64, device=device)
        _lib.call("triplet_count", edge_ptr, member, N, E, nti, ntij, tri_count, member_list, used_count, stats)
        tri_ptr = torch.empty(E + 1, **i32)
        work = torch.empty(_lib.scan_work_elems(E), **i32)
        _lib.call("exclusive_scan_i32", tri_count, tri_ptr, E, work)
        used_ptr = torch.empty(N + 1, **i32)
        work_n = torch.empty(_lib.scan_work_elems(N), **i32)
        _lib.call("exclusive_scan_i32", used_count, used_ptr, N, work_n)
        # one read-back: triplet count, largest member degree, number of bonds that head a triplet
        T, max_n3, n_used = (int(v) for v in stats.tolist())
        tri_e2 = torch.empty(T, **i32)
        tri_index = torch.empty((2, T), dtype=torch.int64, device=device) if want_triplet_index else None
        member_edges = torch.empty(n_used, **i32)
        _lib.call("triplet_fill", edge_ptr, tri_ptr, tri_count, member_list, N, T, tri_e2, tri_index, used_ptr,
                  member_edges)
        g = cls(
            pos=cart64.to(torch.float32), atom_types=types,
            num_triplet_i=nti, edge_index=edge_index, edge_cell_shift=shift, num_triplet_ij=ntij,
            triplet_edge_index=tri_index, lattice=lat64.to(torch.float32) if lat32 is None else lat32,
        )
        g._store[K.BATCH] = batch
        g._store[K.NUM_TRIPLETS] = T
        g._private["edge_distances_build"] = dist
        # the builder already holds the canonical CSR: seed the plan so the model does not re-derive it
        # the builder emits the full off-diagonal pair matrix of every atom's member bonds (the canonical dense layout),
        # so the plan needs no layout check; the largest member count and the member-bond list come from the builder
        plan = GraphPlan.from_builder(g, atom_ptr, edge_ptr, tri_ptr, tri_e2, max_members=max_n3,
                                      member_edges=member_edges)
        plan._type_range = type_range
        object.__setattr__(g, "_plan", plan)
        return g

    @staticmethod
    def _check_device(device):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("graph construction runs on the GPU; no CUDA device is available "
                               "(torch_m3gnet_b200 has no CPU fallback)")
        return torch.device(device)

    @classmethod
    def from_arrays(cls, lattices: np.ndarray, cart, atomic_numbers, sizes: Sequence[int],
                    cutoff: float, threebody_cutoff: float, device: Optional[torch.device] = None,
                    want_triplet_index: bool = True) -> "Batch":
        """Graph build on the GPU from plain arrays: ``lattices`` (B,3,3) rows = cell vectors, ``cart`` (N,3) Cartesian
        coordinates (float64), ``atomic_numbers`` (N), ``sizes`` atoms per structure.  ``cart`` / ``atomic_numbers`` may be
        numpy arrays or torch CPU tensors; pinned CPU tensors are uploaded asynchronously on the current stream."""
        if threebody_cutoff > cutoff:
            raise ValueError("Three body cutoff raidus should be smaller than two body.")
        device = cls._check_device(device)
        with torch.cuda.device(device):
            B = len(sizes)
            N = int(sum(sizes))
            lattices = np.ascontiguousarray(np.asarray(lattices, dtype=np.float64)).reshape(B, 3, 3)
            lat64 = torch.as_tensor(lattices).to(device)
            if torch.is_tensor(cart):
                cart64 = cart.reshape(N, 3).to(device=device, dtype=torch.float64, non_blocking=True)
            else:
                cart64 = torch.as_tensor(np.ascontiguousarray(cart, dtype=np.float64).reshape(N, 3)).to(device)
            atom_ptr_h = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
            atom_ptr = torch.as_tensor(atom_ptr_h).to(device)
            edge_ptr, E, edge_index, shift, dist, member = cls._sweep(lattices, lat64, cart64, atom_ptr, B, N, cutoff,
                                                                      threebody_cutoff, device)
            batch = torch.repeat_interleave(torch.arange(B, device=device), torch.as_tensor(list(sizes), device=device),
                                            output_size=N)  # (output_size: no read-back of the total)
            if torch.is_tensor(atomic_numbers):
                z_h = atomic_numbers.reshape(-1)
                types = (z_h.to(device=device, dtype=torch.int64, non_blocking=True) - 1)
                type_range = (int(z_h.min()) - 1, int(z_h.max()) - 1) if N > 0 else (0, 0)
            else:
                types_h = np.asarray(atomic_numbers, dtype=np.int64) - 1
                types = torch.as_tensor(types_h).to(device)
                type_range = (int(types_h.min()), int(types_h.max())) if N > 0 else (0, 0)
            return cls._assemble(lat64, cart64, types, atom_ptr, batch, N, edge_ptr, E, edge_index, shift, dist,
                                 member, want_triplet_index, type_range=type_range)


class EdgesNotGrouped(ValueError):
    """edge_index[0] is not non-decreasing: the kernels need the bonds of an atom to be contiguous.  ``Gradient``
    catches this and evaluates a regrouped copy of the graph (``regroup_by_source``)."""


def regroup_by_source(g: "MaterialGraph"):
    """Stable sort of the bonds by source atom (ties keep the caller's order): returns (shadow graph whose bond-level
    tensors are regrouped and whose triplet list is renumbered accordingly, ``rank`` with rank[e] = row of the caller's
    bond e in the shadow).  Node- and structure-level tensors are shared with ``g``."""
    ei = g[K.EDGE_INDEX].contiguous()
    dev = ei.device
    E, N = int(ei.size(1)), int(g[K.POS].size(0))
    i32 = dict(dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        src = torch.empty(E, **i32)
        _lib.call("narrow_i64", ei[0].contiguous(), src, E)
        ptr = torch.empty(N + 1, **i32)
        perm = torch.empty(max(E, 1), **i32)
        work = torch.empty(N + 1 + _lib.scan_work_elems(N), **i32)
        _lib.call("csr_by_key", src, E, N, ptr, perm, work)  # rows ascending in the original bond id: stable
    perm = perm[:E].long()
    rank = torch.empty(E, dtype=torch.long, device=dev)
    rank[perm] = torch.arange(E, device=dev)
    tri = g[K.TRIPLET_EDGE_INDEX]
    ntij = g[K.NUM_TRIPLET_IJ]
    shadow = Batch(pos=g[K.POS], atom_types=g[K.ATOM_TYPES], num_triplet_i=g[K.NUM_TRIPLET_I],
                   edge_index=ei[:, perm].contiguous(), edge_cell_shift=g[K.EDGE_CELL_SHIFT][perm].contiguous(),
                   num_triplet_ij=None if ntij is None else ntij[perm].contiguous(),
                   triplet_edge_index=None if tri is None else rank[tri].contiguous(), lattice=g[K.LATTICE])
    if g._store.get(K.BATCH) is not None:
        shadow._store[K.BATCH] = g[K.BATCH]
    return shadow, rank


def _sig(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype, t.device)


class GraphPlan:
    """Canonical int32 views of a batch for the kernels (see include/m3gnet_b200.h "Conventions")."""

    def __init__(self):
        self.N = self.E = self.T = self.B = 0
        self.signature = None
        self.tri_group = 8
        self._type_range = None

    def check_types(self, limit: int, what: str):
        """Raise unless every atom type (= Z - 1) indexes a table of ``limit`` rows.  The reference fails loudly here
        too (nn/featurizer.py:36 one_hot: "Class values must be smaller than num_classes"); the kernels index their
        tables unchecked.  The (min, max) pair is read back once per plan (seeded from the host arrays by the graph
        builder)."""
        if self._type_range is None:
            if self.N == 0:
                self._type_range = (0, 0)
            else:
                lo, hi = torch.aminmax(self.types)
                self._type_range = (int(lo.item()), int(hi.item()))
        lo, hi = self._type_range
        if lo < 0 or hi >= limit:
            raise ValueError(f"atom_types (atomic number - 1) span [{lo}, {hi}] but {what} has {limit} entries "
                             f"(valid types are 0 .. {limit - 1})")

    @staticmethod
    def signature_of(g: MaterialGraph):
        return tuple(_sig(g._store.get(k)) for k in _STRUCTURAL)

    @classmethod
    def from_builder(cls, g, atom_ptr, edge_ptr, tri_ptr, tri_e2, max_members: Optional[int] = None,
                     member_edges: Optional[torch.Tensor] = None) -> "GraphPlan":
        p = cls()
        p._common(g)
        p.atom_ptr, p.edge_ptr = atom_ptr, edge_ptr
        p._in_csr()
        p.tri_ptr, p.tri_e2 = tri_ptr, tri_e2
        p.T = int(tri_e2.numel())  # also when the (2,T) int64 API list was not materialised
        p.trt_ptr, p.trt_e1 = tri_ptr, tri_e2  # builder output is the full off-diagonal: symmetric
        p.tri_symmetric = True
        p._pick_group(dense_max_members=max_members, member_edges=member_edges)
        p.signature = cls.signature_of(g)
        return p

    def _common(self, g):
        ei = g[K.EDGE_INDEX]
        dev = ei.device
        self.device = dev
        self.N = int(g[K.POS].size(0))
        self.E = int(ei.size(1))
        ti = g[K.TRIPLET_EDGE_INDEX]
        self.T = int(ti.size(1)) if ti is not None else 0
        lat = g[K.LATTICE]
        self.B = int(lat.size(0)) if lat.dim() == 3 else 1
        i32 = dict(dtype=torch.int32, device=dev)
        ei = ei.contiguous()
        both = torch.empty((2, self.E), **i32)
        _lib.call("narrow_i64", ei, both, 2 * self.E)
        self.src, self.dst = both[0], both[1]
        self.shift = g[K.EDGE_CELL_SHIFT].to(torch.int32).contiguous()
        self.types = torch.empty(self.N, **i32)
        _lib.call("narrow_i64", g[K.ATOM_TYPES].contiguous(), self.types, self.N)
        batch = g._store.get(K.BATCH)
        if batch is None:
            batch = torch.zeros(self.N, dtype=torch.long, device=dev)
        self.batch = torch.empty(self.N, **i32)
        _lib.call("narrow_i64", batch.contiguous(), self.batch, self.N)

    def _in_csr(self):
        i32 = dict(dtype=torch.int32, device=self.device)
        self.in_ptr = torch.empty(self.N + 1, **i32)
        self.in_perm = torch.empty(max(self.E, 1), **i32)
        work = torch.empty(self.N + 1 + _lib.scan_work_elems(self.N), **i32)
        _lib.call("csr_by_key", self.dst, self.E, self.N, self.in_ptr, self.in_perm, work)

    def _pick_group(self, dense_max_members: Optional[int] = None, member_edges: Optional[torch.Tensor] = None):
        avg = self.T / max(self.E, 1)
        self.tri_group = 8 if avg <= 12 else (16 if avg <= 28 else 32)
        # bonds that are the first bond of at least one triplet ("member" bonds): the only ones whose Bessel basis
        # is ever read
        if member_edges is None:
            used = self.tri_ptr[1:] > self.tri_ptr[:-1]
            if self.trt_ptr is not self.tri_ptr:  # non-symmetric list: bonds that only occur as second bond count too
                used = used | (self.trt_ptr[1:] > self.trt_ptr[:-1])
            member_edges = torch.nonzero(used).flatten().to(torch.int32)
        self.member_edges = member_edges  # (the builder hands over its own list: no compaction pass, no read-back)
        self.n_members = int(self.member_edges.numel())
        # canonical per-atom layout (full off-diagonal of the member-bond pair matrix)?  -> per-atom kernels
        if dense_max_members is not None:  # certified by the builder
            dense, self.max_members = True, int(dense_max_members)
        else:
            flags = torch.empty(2, dtype=torch.int32, device=self.device)
            _lib.call("tri_dense_check", self.src, self.edge_ptr, self.tri_ptr, self.tri_e2, self.E, flags)
            dense, self.max_members = flags.tolist()
        self.tri_dense = bool(dense) and self.max_members <= _lib.tb_atom_capacity()
        self.tri_moment = bool(dense) and self.max_members <= _lib.tb_mom_capacity()

    @classmethod
    def build(cls, g: MaterialGraph) -> "GraphPlan":
        """Canonicalise an arbitrary (possibly hand-built / permuted) graph; raises on unsupported layouts."""
        p = cls()
        p._common(g)
        dev = p.device
        i32 = dict(dtype=torch.int32, device=dev)
        flags = torch.empty(4, **i32)
        # edges must be grouped by source atom (true for every graph produced by from_structure)
        _lib.call("check_sorted", p.src, p.E, p.N, flags)
        f_src = flags[:2].tolist()
        _lib.call("check_sorted", p.dst, p.E, p.N, flags)
        f_dst = flags[:2].tolist()
        if not (f_src[1] and f_dst[1]):
            raise ValueError("edge_index contains atom indices outside [0, num_nodes)")
        if not f_src[0]:
            raise EdgesNotGrouped("edge_index[0] must be non-decreasing (edges grouped by source atom, as "
                                  "MaterialGraph.from_structure produces them); the full model regroups such graphs "
                                  "itself, single layers need regroup_by_source() first")
        _lib.call("check_sorted", p.batch, p.N, p.B, flags)
        f_b = flags[:2].tolist()
        if not (f_b[0] and f_b[1]):
            raise ValueError("batch must be non-decreasing with values in [0, num_structures)")
        p.edge_ptr = torch.empty(p.N + 1, **i32)
        _lib.call("csr_from_sorted", p.src, p.E, p.N, p.edge_ptr)
        p.atom_ptr = torch.empty(p.B + 1, **i32)
        _lib.call("csr_from_sorted", p.batch, p.N, p.B, p.atom_ptr)
        p._in_csr()
        # triplets: CSR per first bond, rows ascending in the second bond
        ti = g[K.TRIPLET_EDGE_INDEX].contiguous()
        both = torch.empty((2, p.T), **i32)
        _lib.call("narrow_i64", ti, both, 2 * p.T)
        e1, e2 = both[0], both[1]
        _lib.call("check_sorted", e1, p.T, p.E, flags)
        f1 = flags[:2].tolist()
        _lib.call("check_sorted", e2, p.T, p.E, flags)
        f2 = flags[:2].tolist()
        if not (f1[1] and f2[1]):
            raise ValueError("triplet_edge_index contains edge indices outside [0, num_edges)")
        p.tri_ptr = torch.empty(p.E + 1, **i32)
        if f1[0]:
            _lib.call("csr_from_sorted", e1, p.T, p.E, p.tri_ptr)
            p.tri_e2 = e2.clone()
        else:
            perm = torch.empty(max(p.T, 1), **i32)
            work = torch.empty(p.E + 1 + _lib.scan_work_elems(p.E), **i32)
            _lib.call("csr_by_key", e1, p.T, p.E, p.tri_ptr, perm, work)
            p.tri_e2 = torch.empty(max(p.T, 1), **i32)
            _lib.call("gather_i32", e2, perm, p.T, p.tri_e2)
        _lib.call("sort_rows", p.tri_ptr, p.E, p.tri_e2)
        _lib.call("csr_is_symmetric", p.tri_ptr, p.tri_e2, p.E, flags)
        p.tri_symmetric = bool(flags[0].item())
        if p.tri_symmetric:
            p.trt_ptr, p.trt_e1 = p.tri_ptr, p.tri_e2
        else:
            p.trt_ptr = torch.empty(p.E + 1, **i32)
            perm = torch.empty(max(p.T, 1), **i32)
            work = torch.empty(p.E + 1 + _lib.scan_work_elems(p.E), **i32)
            _lib.call("csr_by_key", e2, p.T, p.E, p.trt_ptr, perm, work)
            p.trt_e1 = torch.empty(max(p.T, 1), **i32)
            _lib.call("gather_i32", e1, perm, p.T, p.trt_e1)
            _lib.call("sort_rows", p.trt_ptr, p.E, p.trt_e1)
        p._pick_group()
        p.signature = cls.signature_of(g)
        return p


def get_plan(g: MaterialGraph) -> GraphPlan:
    plan = g._plan
    if plan is not None and plan.signature == GraphPlan.signature_of(g):
        return plan
    if not g[K.POS].is_cuda:
        raise RuntimeError("torch_m3gnet_b200 runs on CUDA tensors only (move the batch with .to('cuda')); "
                           "there is no CPU fallback")
    with torch.cuda.device(g[K.POS].device):
        plan = GraphPlan.build(g)
    object.__setattr__(g, "_plan", plan)
    return plan


BatchMaterialGraph = Batch
