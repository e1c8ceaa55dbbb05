"""CPU: the O(n3) moment form of ThreeBodyInteration (oracle/threebody_moments.py, the algorithm of
csrc/threebody_moment.cu) against the oracle's explicit triplet sum + autograd on the live-reference fixture
(tests/golden/threebody_op.npz: O(1) factor table, non-unit upstream gradients, Legendre quirk Q3)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import m3gnet_oracle as O
from oracle import threebody_moments as M
from tests.util import golden, graph_dict, state_dict_of


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-13), (torch.float32, 5e-6)])
def test_moment_form_equals_triplet_sum(dtype, tol):
    g = golden("threebody_op")
    gd = graph_dict(g)
    sd = {"tb." + k: v.to(dtype) for k, v in state_dict_of(g).items()}
    hp = O.HyperParams()
    x = torch.from_numpy(g["x"]).to(dtype).requires_grad_(True)
    e = torch.from_numpy(g["e"]).to(dtype).requires_grad_(True)
    vec = torch.from_numpy(g["vec"]).to(dtype).requires_grad_(True)
    fac = torch.from_numpy(g["factors"]).to(dtype)
    go = torch.from_numpy(g["go"]).to(dtype)
    dist = torch.linalg.norm(vec, dim=1)
    t = gd["triplet_edge_index"]
    cos = torch.clamp(torch.sum(vec[t[0]] * vec[t[1]], dim=1) / (dist[t[0]] * dist[t[1]]), -1, 1)
    out, red = O.three_body(sd, "tb", hp, x, e, dist, cos, gd["edge_index"], t, fac)
    gx_o, gv_o = torch.autograd.grad(out, [x, vec], grad_outputs=go)
    if dtype == torch.float32:  # the fixture holds the live reference's own values
        np.testing.assert_allclose(red.detach().numpy(), g["red"], rtol=2e-5, atol=2e-7)

    L = NM = 3
    rc, r3 = hp.scaled_cutoff, hp.scaled_threebody_cutoff
    src, dst = gd["edge_index"]
    E, N = src.numel(), x.shape[0]
    member = torch.zeros(E, dtype=torch.bool)
    member[t[0]] = True
    vec2 = vec.detach().clone().requires_grad_(True)
    r = torch.linalg.norm(vec2, dim=1)
    u = vec2 / r[:, None]
    c = O.cutoff_function(r, r3)
    zeros = torch.tensor(O.bessel_zero_table().tolist()).to(dtype)
    jl = torch.stack([O.spherical_bessel(zeros[l][:NM, None] * r[None, :] / rc, l) for l in range(L)])
    G = (jl / fac[:, :, None] * c[None, None, :]).permute(2, 0, 1)  # (E, L, NM): the kernel's block-invariant table
    x2 = x.detach().clone().requires_grad_(True)
    sig = torch.sigmoid(F.linear(x2, sd["tb.linear_sigmoid1.weight"], sd["tb.linear_sigmoid1.bias"])).reshape(N, L, NM)
    b = G * sig[dst]
    ud, cd, bd, rd = u.detach(), c.detach(), b.detach(), r.detach()
    red_m = torch.zeros(E, 3, 3, dtype=dtype)
    fwd = {}
    for i in range(N):
        idx = torch.nonzero((src == i) & member).flatten()
        if idx.numel():
            fwd[i] = (idx, M.forward_atom(ud[idx], cd[idx], bd[idx]))
            red_m[idx] = fwd[i][1]["red"]
    scale = red.abs().max().item()
    assert (red_m.reshape(E, 9) - red.detach()).abs().max().item() <= tol * scale
    redl = red_m.reshape(E, 9).clone().requires_grad_(True)
    (q,) = torch.autograd.grad(O.gated_mlp(sd, "tb.gated_mlp", redl, 1, bias=False), redl, grad_outputs=go)
    q = q.reshape(E, 3, 3)
    g_b = torch.zeros(E, 3, 3, dtype=dtype)
    g_c = torch.zeros(E, dtype=dtype)
    g_v = torch.zeros(E, 3, dtype=dtype)
    g_r = torch.zeros(E, dtype=dtype)
    for i, (idx, f) in fwd.items():
        g_b[idx], g_c[idx], g_v[idx], g_r[idx] = M.backward_atom(ud[idx], rd[idx], cd[idx], bd[idx], q[idx], f)
    (gx_m,) = torch.autograd.grad(b, x2, grad_outputs=g_b, retain_graph=True)
    (gvec_rc,) = torch.autograd.grad([b, c], vec2, grad_outputs=[g_b, g_c])
    gvec_m = gvec_rc + g_v + g_r[:, None] * ud
    assert (gx_m - gx_o).abs().max().item() <= tol * gx_o.abs().max().item()
    assert (gvec_m - gv_o).abs().max().item() <= tol * gv_o.abs().max().item()
