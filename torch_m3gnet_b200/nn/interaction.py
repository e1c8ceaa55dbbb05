from __future__ import annotations

import math
import os

import torch

from torch_m3gnet_b200._lib import call
from torch_m3gnet_b200.data import MaterialGraphKey as K
from torch_m3gnet_b200.data.material_graph import get_plan
from torch_m3gnet_b200.nn._bessel_zeros import SPHERICAL_BESSEL_ZEROS
from torch_m3gnet_b200.nn._functions import CutoffFn, GeometryFn, LegendreCosFn, SphericalBesselFn, ThreeBodyFn
from torch_m3gnet_b200.nn._packing import PackedWeights, c_, t_
from torch_m3gnet_b200.nn.core import GatedMLP
from torch_m3gnet_b200.nn.invariant import PAIR_VEC4

# "moment" (default): O(n3)-per-atom moment kernels for (l_max, n_max, F) = (3, 3, 64) when the plan certifies the
# canonical triplet layout (csrc/threebody_moment.cu), falling back like "atom"; "atom": per-centre-atom pair-matrix
# kernels (csrc/threebody_atom.cu), otherwise "fast"; "fast": specialised CSR kernels for (3, 3, 64); "generic": the
# width-agnostic CSR kernels everywhere
TB_PATH = os.environ.get("M3G_TB_PATH", "moment")
RADIAL_CACHE = "_tb_radial"
# moment path, forward: True (default) = per-atom moment kernel (red only) + a streaming edge-update kernel over all bond
# rows; False = one fused per-atom kernel
TB_SPLIT = os.environ.get("M3G_TB_SPLIT", "1") != "0"
# moment path, backward: True (default) = gated-MLP adjoint over the packed member-bond list (a lane owns a row) + the
# per-atom moment kernel reading q = dL/dred; False = one fused per-atom kernel
TB_BWD_SPLIT = os.environ.get("M3G_TB_BWD_SPLIT", "1") != "0"

__all__ = ["ThreeBodyInteration", "NormalizedSphericalBessel", "SPHERICAL_BESSEL_ZEROS", "spherical_bessel",
           "legendre_cos", "cutoff_function"]

# Operator API of the reference (nn/interaction.py:385-400), backed by elementwise CUDA kernels (float32, CUDA
# tensors only).
spherical_bessel = SphericalBesselFn.apply
legendre_cos = LegendreCosFn.apply


def cutoff_function(input: torch.Tensor, cutoff: float) -> torch.Tensor:
    return CutoffFn.apply(input, cutoff)


def _host_bessel_at(x: torch.Tensor, order: int) -> torch.Tensor:
    """j_order(x) on the host with the float32 op sequence of the reference's SphericalBessel.forward
    (nn/interaction.py:288-323).  Used ONLY at construction time for the normalisation table below."""
    eps = 1e-8
    vals = [torch.where(x > eps, torch.sin(x) / x, torch.ones_like(x))]
    if order >= 1:
        vals.append(torch.where(x > eps, (torch.sin(x) / x - torch.cos(x)) / x, x / 3))
        c = 3
        for n in range(1, order):
            c *= 2 * n + 3
            vals.append(torch.where(x > eps, (2 * n + 1) / x * vals[n] - vals[n - 1], x / c))
    return vals[order]


class NormalizedSphericalBessel(torch.nn.Module):
    """Constant tables of the normalised spherical Bessel radial functions (reference nn/interaction.py:226-281).

    ``factors`` reproduces the reference bit-for-bit *on the host*: the reference evaluates j_{l+1} at the zeros
    of j_{l+1} in CPU float32 and divides by it (quirk Q1), so the table is round-off noise of the host's
    sin/cos; it must therefore be computed by the same CPU float32 sequence at construction, never on the GPU.
    Like the reference it is a plain attribute (not a buffer) and may be overwritten by the user."""

    def __init__(self, cutoff: float, l_max: int, n_max: int, device: torch.device | None = None):
        super().__init__()
        self.cutoff = cutoff
        self.l_max = l_max
        self.n_max = n_max
        self.device = device
        zeros = torch.tensor(SPHERICAL_BESSEL_ZEROS)  # float32, as in the reference
        if zeros.size(0) < l_max + 1:
            raise ValueError("Too large l_max is specified.")
        if zeros.size(1) < n_max:
            raise ValueError("Too large n_max is specified.")
        self.spherical_bessel_zeros = zeros.to(device)
        self.factors = torch.stack([
            math.sqrt(2 / (cutoff**3)) / torch.abs(_host_bessel_at(zeros[l + 1, :n_max], l + 1))
            for l in range(l_max)
        ]).to(device)


class ThreeBodyInteration(torch.nn.Module):
    """Three-body update of the edge features (reference nn/interaction.py:138-223), fused on the GPU: the
    triplet gather, basis product and segmented sum into bond features run in one kernel (csrc/threebody_atom.cu,
    csrc/threebody.cu).
    Updates EDGE_ATTR."""

    def __init__(self, cutoff: float, threebody_cutoff: float, l_max: int, n_max: int, num_node_features: int,
                 num_edge_features: int, device: torch.device | None = None):
        super().__init__()
        self.cutoff = cutoff
        self.threebody_cutoff = threebody_cutoff
        self.l_max = l_max
        self.n_max = n_max
        self.degree = l_max * n_max
        self.num_node_features = num_node_features
        self.num_edge_features = num_edge_features
        self.device = device
        self.nsb = NormalizedSphericalBessel(cutoff=cutoff, l_max=l_max, n_max=n_max, device=device)
        self.linear_sigmoid1 = torch.nn.Linear(num_node_features, self.degree, device=device)
        self.gated_mlp = GatedMLP(in_features=self.degree, dimensions=[num_edge_features], use_bias=False,
                                  device=device)
        self._packed = PackedWeights(self._sources, self._pack)

    def _sources(self):
        return [self.linear_sigmoid1.weight, self.linear_sigmoid1.bias, self.gated_mlp.dense[0].weight,
                self.gated_mlp.gate[0].weight, self.nsb.factors, self.nsb.spherical_bessel_zeros]

    def _pack(self):
        dev = self.linear_sigmoid1.weight.device
        L, R = self.l_max, self.n_max
        zeros = self.nsb.spherical_bessel_zeros[:L, :R].to(device=dev, dtype=torch.float32).reshape(-1)
        fac = self.nsb.factors.to(device=dev, dtype=torch.float32).reshape(-1)
        tail = torch.tensor([self.cutoff, self.threebody_cutoff], dtype=torch.float32, device=dev)
        return {
            "Ws": c_(self.linear_sigmoid1.weight), "bs": c_(self.linear_sigmoid1.bias),
            "WdT": t_(self.gated_mlp.dense[0].weight), "WgT": t_(self.gated_mlp.gate[0].weight),
            "consts": torch.cat([zeros, fac, tail]).contiguous(), "r3": float(self.threebody_cutoff),
            # host copy of the constants: key of the per-step cache of the block-invariant radial tables
            "consts_key": tuple(torch.cat([zeros, fac, tail]).tolist()),
        }

    def _radial(self, graph, plan, vec4, w):
        """(G, dG) = m3g_tb_radial for this block's constants; shared by all blocks with equal constants."""
        cache = graph._private.get(RADIAL_CACHE)
        if cache is None or cache["vec4"] is not vec4 or cache["version"] != vec4._version:
            cache = graph._private[RADIAL_CACHE] = {"vec4": vec4, "version": vec4._version, "tables": {}}
        tab = cache["tables"].get(w["consts_key"])
        if tab is None:
            E, D = plan.E, self.degree
            G = torch.empty((E, D), dtype=torch.float32, device=vec4.device)
            dG = torch.empty((E, D), dtype=torch.float32, device=vec4.device)
            call("tb_radial", vec4.detach(), w["consts"], E, self.l_max, self.n_max, plan.member_edges, plan.n_members,
                 G, dG)
            tab = cache["tables"][w["consts_key"]] = (G, dG)
        return tab

    def forward(self, graph):
        plan = get_plan(graph)
        vec4 = graph._private.get(PAIR_VEC4)
        if vec4 is None:
            # called on its own (reference nn/interaction.py:187-192 reads EDGE_DISTANCES / TRIPLET_ANGLES): derive the
            # bond vectors from the public position / lattice keys, as DistanceAndAngle does
            pos = graph[K.SCALED_POS] if graph[K.SCALED_POS] is not None else graph[K.POS]
            lat = graph[K.SCALED_LATTICE] if graph[K.SCALED_LATTICE] is not None else graph[K.LATTICE]
            with torch.cuda.device(pos.device):
                vec4, dist, cos = GeometryFn.apply(pos, lat if lat.dim() == 3 else lat[None], plan,
                                                   graph[K.TRIPLET_EDGE_INDEX])
            graph._private[PAIR_VEC4] = vec4
            if graph[K.EDGE_DISTANCES] is None:
                graph[K.EDGE_DISTANCES] = dist
            if graph[K.TRIPLET_ANGLES] is None:
                graph[K.TRIPLET_ANGLES] = cos
        w = self._packed.get()
        radial = None
        if (TB_PATH == "moment" and (self.l_max, self.n_max, self.num_edge_features) == (3, 3, 64)
                and self.num_node_features == 64 and plan.tri_moment):
            with torch.cuda.device(vec4.device):
                radial = self._radial(graph, plan, vec4, w)
        graph[K.EDGE_ATTR] = ThreeBodyFn.apply(graph[K.NODE_FEATURES], graph[K.EDGE_ATTR], vec4, plan, w, self.l_max,
                                               self.n_max, radial)
        return graph
