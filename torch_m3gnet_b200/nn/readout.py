from __future__ import annotations

import torch

from torch_m3gnet_b200.data import MaterialGraphKey as K
from torch_m3gnet_b200.data.material_graph import get_plan
from torch_m3gnet_b200.nn._functions import ReadoutFn
from torch_m3gnet_b200.nn._packing import PackedWeights, c_, module_params, t_
from torch_m3gnet_b200.nn.core import GatedMLP


class AtomWiseReadout(torch.nn.Module):
    """Atom-wise gated-MLP energies summed per structure (reference nn/readout.py:12-58).  Supplies
    SCALED_ATOMIC_ENERGIES, SCALED_TOTAL_ENERGY, TOTAL_ENERGY."""

    def __init__(self, in_features: int, num_layers: int, scale: float, device: torch.device | None = None):
        super().__init__()
        if num_layers != 3:
            raise ValueError("the fused readout kernel implements the 3-layer readout build_model uses")
        self.in_features = in_features
        self.num_layers = num_layers
        self.scale = scale
        self.gated = GatedMLP(in_features, [in_features] * (num_layers - 1) + [1], is_output=True, device=device)
        self._packed = PackedWeights(module_params(self), self._pack)

    def _pack(self):
        d = self.gated.linears("dense")
        g = self.gated.linears("gate")
        return {
            "W0d": c_(d[0].weight), "W0dT": t_(d[0].weight), "b0d": c_(d[0].bias),
            "W1d": c_(d[1].weight), "W1dT": t_(d[1].weight), "b1d": c_(d[1].bias),
            "w2d": c_(d[2].weight).reshape(-1), "b2d": c_(d[2].bias),
            "W0g": c_(g[0].weight), "W0gT": t_(g[0].weight), "b0g": c_(g[0].bias),
            "W1g": c_(g[1].weight), "W1gT": t_(g[1].weight), "b1g": c_(g[1].bias),
            "w2g": c_(g[2].weight).reshape(-1), "b2g": c_(g[2].bias),
        }

    def forward(self, graph):
        plan = get_plan(graph)
        atomic, stot, tot = ReadoutFn.apply(graph[K.NODE_FEATURES], graph[K.ELEMENTAL_ENERGIES], plan,
                                            self._packed.get(), self.scale)
        graph[K.SCALED_ATOMIC_ENERGIES] = atomic
        graph[K.SCALED_TOTAL_ENERGY] = stot
        graph[K.TOTAL_ENERGY] = tot
        return graph
