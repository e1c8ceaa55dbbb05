from __future__ import annotations

import torch

from torch_m3gnet_b200.data import MaterialGraphKey as K
from torch_m3gnet_b200.data.material_graph import get_plan
from torch_m3gnet_b200.nn._functions import ConvFn
from torch_m3gnet_b200.nn._packing import PackedWeights, c_, t_
from torch_m3gnet_b200.nn.core import GatedMLP


class M3GNetConv(torch.nn.Module):
    """Graph convolution block (reference nn/conv.py:12-97): gated-MLP edge update, then gated-MLP messages
    aggregated on the source atom.  Updates EDGE_ATTR and NODE_FEATURES.

    The first layer acts on cat[x_i, x_j, e]; it is split into per-atom projections of x (one small GEMM per
    block) and a per-edge product with e, so the per-edge work is halved (csrc/conv.cu)."""

    def __init__(self, degree: int, num_node_features: int, num_edge_features: int,
                 device: torch.device | None = None):
        super().__init__()
        if num_node_features != num_edge_features:
            raise ValueError("the fused kernels assume node and edge feature widths are equal "
                             "(build_model always sets num_edge_features = embedding_dim)")
        self.degree = degree
        self.num_node_features = num_node_features
        self.num_edge_features = num_edge_features
        self.num_concat_features = 2 * num_node_features + num_edge_features
        self.concat_edge_update = GatedMLP(self.num_concat_features, [num_edge_features, num_edge_features],
                                           device=device)
        self.edge_linear = torch.nn.Linear(degree, num_edge_features, bias=False, device=device)
        self.concat_node_update = GatedMLP(self.num_concat_features, [num_edge_features, num_node_features],
                                           device=device)
        self.node_linear = torch.nn.Linear(degree, num_node_features, bias=False, device=device)
        self._packed = PackedWeights(lambda: list(self.parameters()), self._pack)

    def _pack_mlp(self, mlp: GatedMLP, lin: torch.nn.Linear):
        F = self.num_node_features
        d0, d1 = mlp.linears("dense")
        g0, g1 = mlp.linears("gate")
        w1e = torch.cat([d0.weight[:, 2 * F:], g0.weight[:, 2 * F:]], dim=0).detach()  # (2F, F) (out,in)
        return {
            "W1e": w1e.contiguous(), "W1eT": w1e.t().contiguous(),
            "W2d": c_(d1.weight), "W2dT": t_(d1.weight), "b2d": c_(d1.bias),
            "W2g": c_(g1.weight), "W2gT": t_(g1.weight), "b2g": c_(g1.bias),
            "Wh": c_(lin.weight), "WhT": t_(lin.weight),
        }, d0, g0

    def _pack(self):
        F = self.num_node_features
        edge, ed0, eg0 = self._pack_mlp(self.concat_edge_update, self.edge_linear)
        node, nd0, ng0 = self._pack_mlp(self.concat_node_update, self.node_linear)
        zeros = torch.zeros(2 * F, dtype=torch.float32, device=ed0.weight.device)
        # per-atom projection: columns [edge: src(dense|gate) dst(dense|gate) | node: same], bias on the src part
        wp = torch.cat([
            ed0.weight[:, :F], eg0.weight[:, :F], ed0.weight[:, F:2 * F], eg0.weight[:, F:2 * F],
            nd0.weight[:, :F], ng0.weight[:, :F], nd0.weight[:, F:2 * F], ng0.weight[:, F:2 * F],
        ], dim=0).detach()  # (8F, F) (out,in)
        bp = torch.cat([ed0.bias, eg0.bias, zeros, nd0.bias, ng0.bias, zeros]).detach()
        return {"edge": edge, "node": node, "Wp": wp.contiguous(), "WpT": wp.t().contiguous(),
                "bp": bp.contiguous()}

    def forward(self, graph):
        plan = get_plan(graph)
        x, e = ConvFn.apply(graph[K.NODE_FEATURES], graph[K.EDGE_ATTR], graph[K.EDGE_WEIGHTS], plan,
                            self._packed.get())
        graph[K.NODE_FEATURES] = x
        graph[K.EDGE_ATTR] = e
        return graph
