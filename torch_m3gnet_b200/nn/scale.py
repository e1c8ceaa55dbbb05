from __future__ import annotations

import torch

from torch_m3gnet_b200.data import MaterialGraphKey as K
from torch_m3gnet_b200.nn._functions import ScaleFn


class ScaleLength(torch.nn.Module):
    """Length-unit normalisation (reference nn/scale.py:9-29): supplies SCALED_POS and SCALED_LATTICE."""

    def __init__(self, length_scale: float):
        super().__init__()
        self.length_scale = length_scale

    def forward(self, graph):
        graph[K.SCALED_POS] = ScaleFn.apply(graph[K.POS], self.length_scale)
        graph[K.SCALED_LATTICE] = ScaleFn.apply(graph[K.LATTICE], self.length_scale)
        return graph
