from __future__ import annotations

import torch

from torch_m3gnet_b200.nn.atom_ref import AtomRef
from torch_m3gnet_b200.nn.conv import M3GNetConv
from torch_m3gnet_b200.nn.featurizer import AtomFeaturizer, EdgeAdjustor, EdgeFeaturizer
from torch_m3gnet_b200.nn.gradient import Gradient
from torch_m3gnet_b200.nn.interaction import ThreeBodyInteration
from torch_m3gnet_b200.nn.invariant import DistanceAndAngle
from torch_m3gnet_b200.nn.readout import AtomWiseReadout
from torch_m3gnet_b200.nn.scale import ScaleLength


def build_model(
    cutoff: float,
    threebody_cutoff: float,
    l_max: int,
    n_max: int,
    num_types: int,
    embedding_dim: int,
    num_blocks: int,
    elemental_energies: torch.Tensor | None = None,
    energy_scale: float = 1.0,  # eV
    length_scale: float = 1.0,  # AA
    device: torch.device | None = None,
) -> torch.nn.Module:
    """Same signature, module order and state_dict keys as the reference (model/build.py:16-83):
    Gradient(Sequential[ScaleLength, AtomRef, DistanceAndAngle, AtomFeaturizer, EdgeFeaturizer, EdgeAdjustor,
    (ThreeBodyInteration, M3GNetConv) x num_blocks, AtomWiseReadout])."""
    if elemental_energies is None:
        elemental_energies = torch.zeros(num_types, device=device)
    rc = cutoff / length_scale
    r3 = threebody_cutoff / length_scale
    layers = [
        ScaleLength(length_scale=length_scale),
        AtomRef(elemental_energies, device=device),
        DistanceAndAngle(),
        AtomFeaturizer(num_types=num_types, embedding_dim=embedding_dim, device=device),
        EdgeFeaturizer(degree=n_max, cutoff=rc, device=device),
        EdgeAdjustor(degree=n_max, num_edge_features=embedding_dim, device=device),
    ]
    for _ in range(num_blocks):
        layers.append(ThreeBodyInteration(cutoff=rc, threebody_cutoff=r3, l_max=l_max, n_max=n_max,
                                          num_node_features=embedding_dim, num_edge_features=embedding_dim,
                                          device=device))
        layers.append(M3GNetConv(degree=n_max, num_node_features=embedding_dim, num_edge_features=embedding_dim,
                                 device=device))
    layers.append(AtomWiseReadout(in_features=embedding_dim, num_layers=3, scale=energy_scale, device=device))
    return Gradient(torch.nn.Sequential(*layers))
