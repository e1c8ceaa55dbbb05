"""Dev tool: fp32 error of the three-body paths (moment / atom / generic) against the oracle evaluated in FLOAT64 on a
C5-density sub-box.  Prints max abs errors of out, g_x and the total bond-vector gradient per path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import m3gnet_oracle as O  # noqa: E402
from torch_m3gnet_b200 import Batch, synthetic  # noqa: E402
from torch_m3gnet_b200.nn import interaction  # noqa: E402
from torch_m3gnet_b200.nn._functions import GeometryFn  # noqa: E402
from torch_m3gnet_b200.nn.invariant import PAIR_VEC4  # noqa: E402

dev = torch.device("cuda:0")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lat, cart, z = synthetic.fcc_cu_supercell(cells, 0.4, 5)
b = Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 5.0, device=dev)
plan = b._plan
tb = interaction.ThreeBodyInteration(5.0, 5.0, 3, 3, 64, 64, device=dev)
fac = torch.rand(3, 3, generator=torch.Generator().manual_seed(3)) + 0.5
tb.nsb.factors = fac.to(dev)
vec4, dist, cos = GeometryFn.apply(b["pos"], b["lattice"], plan, b["triplet_edge_index"])
vec4 = vec4.detach()
torch.manual_seed(5)
x0 = 0.1 * torch.randn(plan.N, 64)
e0 = 0.05 * torch.randn(plan.E, 64)
go = torch.randn(plan.E, 64)
# float64 oracle with the bond vectors as leaves (same fp32 vectors, promoted)
hp = O.HyperParams(threebody_cutoff=5.0)
sd = {"tb." + k: v.detach().cpu().double() for k, v in tb.state_dict().items()}
vec = vec4[:, :3].cpu().double().requires_grad_(True)
xd = x0.double().requires_grad_(True)
d64 = torch.linalg.norm(vec, dim=1)
t = b["triplet_edge_index"].cpu()
c64 = torch.clamp((vec[t[0]] * vec[t[1]]).sum(1) / (d64[t[0]] * d64[t[1]]), -1, 1)
out_o, _ = O.three_body(sd, "tb", hp, xd, e0.double(), d64, c64, b["edge_index"].cpu(), t, fac.double())
gx_o, gv_o = torch.autograd.grad(out_o, [xd, vec], grad_outputs=go.double())
print(f"atoms {plan.N} bonds {plan.E} triplets {plan.T} max_members {plan.max_members}; "
      f"max|out| {out_o.abs().max():.3e} max|g_x| {gx_o.abs().max():.3e} max|g_vec| {gv_o.abs().max():.3e}")
for path in ("moment", "atom", "generic"):
    interaction.TB_PATH = path
    x = x0.to(dev).requires_grad_(True)
    v4 = vec4.clone().requires_grad_(True)
    b._private.clear()
    b._private[PAIR_VEC4] = v4
    b["x"], b["edge_attr"] = x, e0.to(dev)
    out = tb(b)["edge_attr"]
    gx, gv = torch.autograd.grad(out, [x, v4], grad_outputs=go.to(dev))
    gvec = (gv[:, :3] + gv[:, 3:4] * vec4[:, :3] / vec4[:, 3:4]).cpu().double()
    err = (gvec - gv_o).abs()
    worst = int(err.max(dim=1).values.argmax())
    print(f"{path:8s} out {(out.cpu().double() - out_o).abs().max():.3e}  g_x {(gx.cpu().double() - gx_o).abs().max():.3e}  "
          f"g_vec {err.max():.3e} (bond {worst}, r={vec4[worst, 3].item():.4f}, |gv4|={gv[worst].abs().max().item():.3e})")
