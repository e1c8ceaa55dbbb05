from __future__ import annotations

import numpy as np
import torch

from torch_m3gnet_b200._lib import call
from torch_m3gnet_b200.data import MaterialGraphKey as K
from torch_m3gnet_b200.data.material_graph import get_plan
from torch_m3gnet_b200.nn._functions import EdgeAdjustFn, RadialFn
from torch_m3gnet_b200.nn._packing import PackedWeights, c_, t_


class AtomFeaturizer(torch.nn.Module):
    """x = W_emb[:, Z-1] (reference nn/featurizer.py:11-38 does one_hot · Linear; a column gather is the same
    arithmetic).  Supplies NODE_FEATURES."""

    def __init__(self, num_types: int, embedding_dim: int, device: torch.device | None = None):
        super().__init__()
        self._num_types = num_types
        self.linear = torch.nn.Linear(num_types, embedding_dim, bias=False, device=device)
        self._packed = PackedWeights(lambda: [self.linear.weight], lambda: {"W": c_(self.linear.weight)})

    @property
    def num_types(self) -> int:
        return self._num_types

    def forward(self, graph):
        plan = get_plan(graph)
        w = self._packed.get()["W"]
        plan.check_types(self._num_types, "the atom embedding (num_types)")
        F = w.shape[0]
        x = torch.empty((plan.N, F), dtype=torch.float32, device=plan.device)
        call("embed_fwd", w, plan.types, plan.N, F, self._num_types, x)
        graph[K.NODE_FEATURES] = x
        return graph


def radial_constants(degree: int, cutoff: float) -> torch.Tensor:
    """Host-side constant table of the radial basis, produced with the reference's float32 op sequence
    (nn/featurizer.py:61-79, 86-96): [k_0..k_R | coeff | sqrt(e_m/d_{m-1}) | sqrt(d_m)]."""
    iota = torch.arange(degree)
    em = (iota**2) * ((iota + 2) ** 2) / (4 * ((iota + 1) ** 4) + 1)
    dm = torch.ones(degree)
    for m in range(1, degree):
        dm[m] = 1 - em[m] / dm[m - 1]
    coeff = torch.empty(degree)
    for m in range(degree):
        coeff[m] = (((-1) ** m) * np.sqrt(2) * np.pi / (cutoff**1.5) * (m + 1) * (m + 2)
                    / np.sqrt((m + 1) ** 2 + (m + 2) ** 2))
    k = (torch.arange(degree + 1) + 1) * torch.pi / cutoff  # float32, as the reference forms sinc's argument
    a = torch.zeros(degree)
    for m in range(1, degree):
        a[m] = torch.sqrt(em[m] / dm[m - 1])
    b = torch.sqrt(dm)
    return torch.cat([k.to(torch.float32), coeff, a, b]).contiguous(), em, dm, coeff


class EdgeFeaturizer(torch.nn.Module):
    """Orthogonalised smooth radial basis h_m(r), m < degree (reference nn/featurizer.py:41-100; note the
    normalised sinc).  Supplies EDGE_WEIGHTS (E, degree)."""

    def __init__(self, degree: int, cutoff: float, device: torch.device | None = None):
        super().__init__()
        self.degree = degree
        self.cutoff = cutoff
        self.device = device
        consts, em, dm, coeff = radial_constants(degree, cutoff)
        self.em, self.dm, self.coeff = em.to(device), dm.to(device), coeff.to(device)
        self._consts_host = consts
        self._consts = {}

    def _device_consts(self, device):
        if device not in self._consts:
            self._consts[device] = self._consts_host.to(device)
        return self._consts[device]

    def forward(self, graph):
        dist = graph[K.EDGE_DISTANCES]
        graph[K.EDGE_WEIGHTS] = RadialFn.apply(dist, self._device_consts(dist.device), self.degree)
        return graph


class EdgeAdjustor(torch.nn.Module):
    """Initial edge features e0 = SiLU(W h) (reference nn/featurizer.py:103-132).  Supplies EDGE_ATTR."""

    def __init__(self, degree: int, num_edge_features: int, device: torch.device | None = None):
        super().__init__()
        self.degree = degree
        self.num_edge_features = num_edge_features
        self.linear = torch.nn.Linear(degree, num_edge_features, bias=False, device=device)
        self.swish = torch.nn.SiLU()
        self._packed = PackedWeights(lambda: [self.linear.weight], lambda: {"Wt": t_(self.linear.weight)})

    def forward(self, graph):
        graph[K.EDGE_ATTR] = EdgeAdjustFn.apply(graph[K.EDGE_WEIGHTS], self._packed.get()["Wt"])
        return graph
