"""Verlet (skin) neighbour list for MD / relaxation loops.

The reference builds every graph from scratch on the host (``MaterialGraph.from_structure``,
torch_m3gnet/data/material_graph.py:132-165, pymatgen neighbour search :168-193 + the Python triplet loop :196-254); a
trajectory pays that for every frame.  ``VerletList`` keeps, on the GPU, a candidate list built once with
``cutoff + skin`` and turns it into the graph of each new frame with one filtering pass:

* the accept test is the builder's own float64 test on the candidates, emitted in the builder's (j, s0, s1, s2) order,
  so the ``Batch`` of a frame is **bit-identical** to ``Batch.from_arrays`` on the same coordinates (same edge_index,
  edge_cell_shift, triplet tensors) -- as long as no atom has moved further than skin/2 since the candidates were built;
* that condition is checked on the GPU every frame (``m3g_verlet_displacement``); when it fails, or when the lattice
  changes, the candidates are rebuilt with the ordinary cell-list sweep.

Coordinates are used as given (never wrapped back into the cell): bond images stay relative to the unwrapped
coordinates, exactly as in the builder.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from torch_m3gnet_b200 import _lib
from torch_m3gnet_b200.data.material_graph import Batch


class VerletList:
    def __init__(self, lattices, atomic_numbers, sizes: Sequence[int], cutoff: float, threebody_cutoff: float,
                 skin: float = 0.5, device: Optional[torch.device] = None, want_triplet_index: bool = False):
        if threebody_cutoff > cutoff:
            raise ValueError("Three body cutoff raidus should be smaller than two body.")
        if not skin > 0.0:
            raise ValueError("skin must be positive")
        self.device = Batch._check_device(device)
        self.cutoff, self.threebody_cutoff, self.skin = float(cutoff), float(threebody_cutoff), float(skin)
        self.want_triplet_index = want_triplet_index
        self.sizes = [int(n) for n in sizes]
        self.B, self.N = len(self.sizes), int(sum(self.sizes))
        with torch.cuda.device(self.device):
            dev = self.device
            self.atom_ptr = torch.as_tensor(np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int32)).to(dev)
            self.batch = torch.repeat_interleave(torch.arange(self.B, device=dev),
                                                 torch.as_tensor(self.sizes, device=dev))
            types_h = np.asarray(atomic_numbers, dtype=np.int64).reshape(-1) - 1
            self.types = torch.as_tensor(types_h).to(dev)
            self.type_range = (int(types_h.min()), int(types_h.max())) if types_h.size else (0, 0)
            if self.types.numel() != self.N:
                raise ValueError("atomic_numbers and sizes disagree")
            self._max_d2 = torch.zeros(1, dtype=torch.float32, device=dev)
        self.set_lattice(lattices)
        self.n_frames = self.n_rebuilds = 0

    def set_lattice(self, lattices):
        """A new cell invalidates the candidates (images and the skin argument are tied to the lattice)."""
        lat = np.ascontiguousarray(np.asarray(lattices, dtype=np.float64)).reshape(self.B, 3, 3)
        self._lattices_h = lat
        with torch.cuda.device(self.device):
            self.lat64 = torch.as_tensor(lat).to(self.device)
            self.lat32 = self.lat64.to(torch.float32)
        self._ref = None

    def _cart64(self, cart) -> torch.Tensor:
        if torch.is_tensor(cart):
            c = cart.to(device=self.device, dtype=torch.float64)
        else:
            c = torch.as_tensor(np.ascontiguousarray(cart, dtype=np.float64)).to(self.device)
        c = c.reshape(self.N, 3).contiguous()
        return c

    def _rebuild(self, cart64):
        r = self.cutoff + self.skin
        ptr, C, index, shift, _, _ = Batch._sweep(self._lattices_h, self.lat64, cart64, self.atom_ptr, self.B, self.N,
                                                  r, r, self.device)
        self.cand_ptr, self.cand_shift, self.C = ptr, shift, C
        self.cand_j = torch.empty(max(C, 1), dtype=torch.int32, device=self.device)
        _lib.call("narrow_i64", index[1].contiguous(), self.cand_j, C)
        self._ref = cart64.clone()
        self.n_rebuilds += 1

    def filter(self, cart):
        """Bond tensors of the frame ``cart`` (candidates rebuilt first when the skin is exhausted): a tuple
        (cart64, edge_ptr, E, edge_index, shift, dist, member) for ``assemble``."""
        with torch.cuda.device(self.device):
            cart64 = self._cart64(cart)
            self.n_frames += 1
            if self._ref is not None:
                _lib.call("verlet_displacement", cart64, self._ref, self.N, self._max_d2)
                # strict: at exactly skin/2 the triangle bound d_old <= d_new + skin still holds, but keep a margin
                # for the float64 rounding of the two distances
                if not float(self._max_d2.item()) < (0.5 * self.skin) ** 2 * (1.0 - 1e-6):
                    self._ref = None
            if self._ref is None:
                self._rebuild(cart64)
            dev, N, B = self.device, self.N, self.B
            i32 = dict(dtype=torch.int32, device=dev)
            counts = torch.empty(N, **i32)
            _lib.call("verlet_count", self.lat64, cart64, self.atom_ptr, B, N, self.cutoff, self.cand_ptr, self.cand_j,
                      self.cand_shift, counts)
            edge_ptr = torch.empty(N + 1, **i32)
            work = torch.empty(_lib.scan_work_elems(N), **i32)
            _lib.call("exclusive_scan_i32", counts, edge_ptr, N, work)
            E = int(edge_ptr[-1].item())
            edge_index = torch.empty((2, E), dtype=torch.int64, device=dev)
            shift = torch.empty((E, 3), **i32)
            dist = torch.empty(E, dtype=torch.float32, device=dev)
            member = torch.empty(E, **i32)
            _lib.call("verlet_fill", self.lat64, cart64, self.atom_ptr, B, N, self.cutoff, self.threebody_cutoff,
                      self.cand_ptr, self.cand_j, self.cand_shift, edge_ptr, E, edge_index, shift, dist, member)
            return cart64, edge_ptr, E, edge_index, shift, dist, member

    def assemble(self, frame) -> Batch:
        """Triplets + Batch (with its plan) of a ``filter`` result."""
        cart64, edge_ptr, E, edge_index, shift, dist, member = frame
        with torch.cuda.device(self.device):
            return Batch._assemble(self.lat64, cart64, self.types, self.atom_ptr, self.batch, self.N, edge_ptr, E,
                                   edge_index, shift, dist, member, self.want_triplet_index, lat32=self.lat32,
                                   type_range=self.type_range)

    def update(self, cart) -> Batch:
        """Graph of the frame with coordinates ``cart`` ((N,3) numpy array or float64 CUDA tensor, Cartesian, A)."""
        return self.assemble(self.filter(cart))

    @staticmethod
    def same_bonds(a, b) -> bool:
        """Do two ``filter`` results hold the same bonds, images and three-body membership (same graph topology:
        identical index tensors, triplets and plan; only coordinates differ)?"""
        return (a[2] == b[2] and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4]) and torch.equal(a[6], b[6]))
