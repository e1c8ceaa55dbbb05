"""A minimal periodic-structure record with the duck type ``MaterialGraph.from_structure`` reads
(``.lattice.matrix``, ``.cart_coords``, iteration over sites with ``.specie.Z``, ``len``) — the subset of
pymatgen's ``Structure`` the reference touches (data/material_graph.py:139-147,177).  pymatgen itself is
not a dependency; a real pymatgen ``Structure`` works unchanged."""
from __future__ import annotations

from typing import Sequence

import numpy as np

_SYMBOLS = (
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
    "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
    "Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr"
).split()
_Z_OF = {s: i + 1 for i, s in enumerate(_SYMBOLS)}


class _Lattice:
    def __init__(self, matrix):
        self.matrix = np.array(matrix, dtype=np.float64).reshape(3, 3)


class _Specie:
    def __init__(self, z: int):
        self.Z = int(z)


class _Site:
    def __init__(self, z: int):
        self.specie = _Specie(z)


class Structure:
    def __init__(self, lattice, species: Sequence, coords, coords_are_cartesian: bool = False):
        self.lattice = _Lattice(lattice)
        coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
        self.cart_coords = coords.copy() if coords_are_cartesian else coords @ self.lattice.matrix
        self.atomic_numbers = np.array([s if isinstance(s, (int, np.integer)) else _Z_OF[s] for s in species],
                                       dtype=np.int64)
        if len(self.atomic_numbers) != len(self.cart_coords):
            raise ValueError("species and coords differ in length")

    @property
    def frac_coords(self):
        return self.cart_coords @ np.linalg.inv(self.lattice.matrix)

    def __len__(self):
        return len(self.cart_coords)

    def __iter__(self):
        return (_Site(z) for z in self.atomic_numbers)
