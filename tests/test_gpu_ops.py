"""GPU: every CUDA operator (through the C ABI) against the oracle / the live-reference fixtures."""
import numpy as np
import pytest
import torch

from oracle import m3gnet_oracle as O
from tests.util import golden, graph_dict, report, state_dict_of, to_batch

pytestmark = pytest.mark.gpu


def _plan_and_geometry(gd, device):
    from torch_m3gnet_b200.data.material_graph import get_plan
    from torch_m3gnet_b200.nn._functions import GeometryFn

    b = to_batch(gd, device)
    plan = get_plan(b)
    pos = b["pos"].clone().requires_grad_(True)
    vec4, dist, cos = GeometryFn.apply(pos, b["lattice"], plan, b["triplet_edge_index"])
    return b, plan, pos, vec4, dist, cos


def test_geometry_fwd_bwd(device):
    gd = graph_dict(golden("small_batch"))
    b, plan, pos, vec4, dist, cos = _plan_and_geometry(gd, device)
    pos_c = gd["pos"].clone().requires_grad_(True)
    vec_o, dist_o, cos_o = O.pair_geometry(pos_c, gd["lattice"], gd["batch"], gd["edge_index"],
                                           gd["edge_cell_shift"], gd["triplet_edge_index"])
    report("geom.vec", vec4[:, :3], vec_o, 1e-7, 1e-6)
    report("geom.dist", dist, dist_o, 1e-7, 1e-6)
    report("geom.cos", cos, cos_o, 5e-7, 0)
    print("[parity] geom.dist bitwise mismatches:", int((dist.cpu() != dist_o.detach()).sum()), "of", dist.numel())
    torch.manual_seed(0)
    gv = torch.randn(vec_o.shape)
    gdist = torch.randn(dist_o.shape)
    gcos = torch.randn(cos_o.shape)
    (want,) = torch.autograd.grad([vec_o, dist_o, cos_o], pos_c, grad_outputs=[gv, gdist, gcos])
    gv4 = torch.cat([gv, torch.zeros(gv.shape[0], 1)], dim=1).to(device)
    (got,) = torch.autograd.grad([vec4, dist, cos], pos, grad_outputs=[gv4, gdist.to(device), gcos.to(device)])
    report("geom.g_pos", got, want, 1e-5, 1e-5)


def test_radial_and_edge_adjust(device):
    from torch_m3gnet_b200.nn._functions import EdgeAdjustFn, RadialFn
    from torch_m3gnet_b200.nn.featurizer import radial_constants

    torch.manual_seed(1)
    r = (torch.rand(999) * 4.9 + 0.1)
    r[0] = 5.0
    for R, rc in ((3, 5.0), (4, 4.2), (1, 5.0)):
        rc_ = r.clone().requires_grad_(True)
        want = O.radial_basis(rc_, R, rc)
        go = torch.randn(want.shape)
        (gw,) = torch.autograd.grad(want, rc_, grad_outputs=go)
        consts = radial_constants(R, rc)[0].to(device)
        rg = r.to(device).requires_grad_(True)
        got = RadialFn.apply(rg, consts, R)
        (gg,) = torch.autograd.grad(got, rg, grad_outputs=go.to(device))
        report(f"radial.h R={R}", got, want, 2e-7, 2e-6)
        report(f"radial.g_r R={R}", gg, gw, 2e-6, 2e-5)
    g = golden("basis")
    consts = radial_constants(3, 5.0)[0].to(device)
    report("radial.known", RadialFn.apply(torch.from_numpy(g["r"]).to(device), consts, 3), g["h"], 2e-7, 2e-6)
    # edge adjustor
    for F in (64, 17):
        W = torch.randn(F, 3) * 0.5
        h = (torch.randn(301, 3) * 0.3).requires_grad_(True)
        want = torch.nn.functional.silu(torch.nn.functional.linear(h, W))
        go = torch.randn(want.shape)
        (gw,) = torch.autograd.grad(want, h, grad_outputs=go)
        hg = h.detach().to(device).requires_grad_(True)
        got = EdgeAdjustFn.apply(hg, W.t().contiguous().to(device))
        (gg,) = torch.autograd.grad(got, hg, grad_outputs=go.to(device))
        report(f"adjust.e0 F={F}", got, want, 1e-6, 1e-6)
        report(f"adjust.g_h F={F}", gg, gw, 1e-5, 1e-5)


def test_basis_operator_api(device):
    from torch_m3gnet_b200.nn.interaction import cutoff_function, legendre_cos, spherical_bessel

    g = golden("basis")
    xs = torch.from_numpy(g["leg_x"]).to(device).requires_grad_(True)
    xb = torch.from_numpy(g["bes_x"]).to(device).requires_grad_(True)
    for l in range(4):
        y = legendre_cos(xs, l)
        (gl,) = torch.autograd.grad(y, xs, grad_outputs=torch.full_like(xs, 0.5))
        report(f"api.leg{l}", y, g[f"leg{l}"], 1e-6, 1e-6)
        report(f"api.leg{l}.grad", gl, g[f"leg{l}_grad_go0.5"], 1e-6, 1e-6)  # quirk Q3
        y = spherical_bessel(xb, l)
        (gl,) = torch.autograd.grad(y, xb, grad_outputs=torch.ones_like(xb))
        report(f"api.j{l}", y, g[f"j{l}"], 2e-6, 1e-5)
        report(f"api.j{l}.grad", gl, g[f"j{l}_grad"], 1e-5, 1e-5)
    report("api.fc", cutoff_function(torch.tensor([1.0, 2.556, 3.615, 4.0, 4.5], device=device), 4.0), g["fc4"],
           1e-7, 1e-6)
    report("api.fc2", cutoff_function(torch.tensor([0.0, 2.0, 4.0], device=device), 2.0), torch.tensor([1.0, 0, 0]),
           0, 0)
    # reference tests/test_basis.py:15-22: j_l(zero) ~ 0
    from torch_m3gnet_b200.nn.interaction import SPHERICAL_BESSEL_ZEROS
    for l in range(len(SPHERICAL_BESSEL_ZEROS)):
        z = torch.tensor(SPHERICAL_BESSEL_ZEROS[l], device=device)
        assert spherical_bessel(z, l).abs().max().item() < 1e-5


def _threebody_reference_grads(g, gd):
    """Oracle autograd with the bond vectors as leaves: returns grads w.r.t. x, e and vec (E,3)."""
    sd = {"tb." + k: v for k, v in state_dict_of(g).items()}
    hp = O.HyperParams()
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    e = torch.from_numpy(g["e"]).requires_grad_(True)
    vec = torch.from_numpy(g["vec"]).requires_grad_(True)
    dist = torch.linalg.norm(vec, dim=1)
    t = gd["triplet_edge_index"]
    cos = torch.clamp(torch.sum(vec[t[0]] * vec[t[1]], dim=1) / (dist[t[0]] * dist[t[1]]), -1, 1)
    out, red = O.three_body(sd, "tb", hp, x, e, dist, cos, gd["edge_index"], t, torch.from_numpy(g["factors"]))
    gx, ge, gv = torch.autograd.grad(out, [x, e, vec], grad_outputs=torch.from_numpy(g["go"]))
    return out, red, gx, ge, gv


@pytest.fixture(params=["moment", "moment-fused", "atom", "fast", "generic"])
def tb_path(request):
    """"moment" = the default split kernels (moments + streaming MLP / MLP adjoint); "moment-fused" = the one-kernel
    per-atom forms of the same path (M3G_TB_SPLIT=0, M3G_TB_BWD_SPLIT=0)."""
    from torch_m3gnet_b200.nn import interaction

    old = interaction.TB_PATH, interaction.TB_SPLIT, interaction.TB_BWD_SPLIT
    interaction.TB_PATH = request.param.split("-")[0]
    if request.param == "moment-fused":
        interaction.TB_SPLIT = interaction.TB_BWD_SPLIT = False
    yield request.param
    interaction.TB_PATH, interaction.TB_SPLIT, interaction.TB_BWD_SPLIT = old


def test_threebody_operator(device, tb_path):
    """Operator-level parity with an O(1) factor table and non-unit upstream gradients (quirks Q1, Q3)."""
    from torch_m3gnet_b200.nn.interaction import ThreeBodyInteration

    g = golden("threebody_op")
    gd = graph_dict(g)
    b, plan, pos, vec4, dist, cos = _plan_and_geometry(gd, device)
    report("tb.geom.dist", dist, g["dist"], 1e-6, 1e-6)
    assert plan.tri_dense and plan.tri_moment, "compute_threebody layout must be certified dense (per-atom kernels would be skipped)"
    tb = ThreeBodyInteration(5.0, 4.0, 3, 3, 64, 64, device=device)
    tb.load_state_dict(state_dict_of(g))
    tb.nsb.factors = torch.from_numpy(g["factors"]).to(device)
    for group in (8, 16, 32):
        plan.tri_group = group
        x = torch.from_numpy(g["x"]).to(device).requires_grad_(True)
        e = torch.from_numpy(g["e"]).to(device).requires_grad_(True)
        v4 = vec4.detach().clone().requires_grad_(True)
        b["x"], b["edge_attr"] = x, e
        b._private["_pair_vec4"] = v4
        out = tb(b)["edge_attr"]
        report(f"tb.out G={group}", out, g["out"], 2e-6, 2e-6)
        gx, ge, gv4 = torch.autograd.grad(out, [x, e, v4], grad_outputs=torch.from_numpy(g["go"]).to(device))
        _, red_o, gx_o, ge_o, gv_o = _threebody_reference_grads(g, gd)
        report(f"tb.gx G={group}", gx, g["gx"], 1e-5, 2e-5)
        report(f"tb.ge G={group}", ge, g["ge"], 0, 0)
        v = v4.detach()
        gv = gv4[:, :3] + gv4[:, 3:4] * v[:, :3] / v[:, 3:4]
        report(f"tb.gvec G={group}", gv, gv_o, 2e-5, 2e-5)
    # end-to-end to positions (through the geometry adjoint), default group
    plan.tri_group = 8
    x = torch.from_numpy(g["x"]).to(device)
    e = torch.from_numpy(g["e"]).to(device)
    b["x"], b["edge_attr"] = x, e
    b._private["_pair_vec4"] = vec4
    out = tb(b)["edge_attr"]
    (gp,) = torch.autograd.grad(out, pos, grad_outputs=torch.from_numpy(g["go"]).to(device))
    pos_c = gd["pos"].clone().requires_grad_(True)
    vec_o, dist_o, cos_o = O.pair_geometry(pos_c, gd["lattice"], gd["batch"], gd["edge_index"],
                                           gd["edge_cell_shift"], gd["triplet_edge_index"])
    sd = {"tb." + k: v for k, v in state_dict_of(g).items()}
    out_o, _ = O.three_body(sd, "tb", O.HyperParams(), torch.from_numpy(g["x"]), torch.from_numpy(g["e"]), dist_o,
                            cos_o, gd["edge_index"], gd["triplet_edge_index"], torch.from_numpy(g["factors"]))
    (gp_o,) = torch.autograd.grad(out_o, pos_c, grad_outputs=torch.from_numpy(g["go"]))
    report("tb.g_pos", gp, gp_o, 5e-5, 5e-5)


def test_threebody_asymmetric_triplet_list(device, tb_path):
    """A hand-made triplet list that is not symmetric exercises the transposed CSR path."""
    from torch_m3gnet_b200.nn.interaction import ThreeBodyInteration

    g = golden("threebody_op")
    gd = graph_dict(g)
    torch.manual_seed(5)
    T = gd["triplet_edge_index"].shape[1]
    keep = torch.rand(T) < 0.6
    gd["triplet_edge_index"] = gd["triplet_edge_index"][:, keep][:, torch.randperm(int(keep.sum()))]
    b, plan, pos, vec4, dist, cos = _plan_and_geometry(gd, device)
    assert not plan.tri_symmetric and not plan.tri_dense
    tb = ThreeBodyInteration(5.0, 4.0, 3, 3, 64, 64, device=device)
    tb.load_state_dict(state_dict_of(g))
    tb.nsb.factors = torch.from_numpy(g["factors"]).to(device)
    b["x"], b["edge_attr"] = torch.from_numpy(g["x"]).to(device), torch.from_numpy(g["e"]).to(device)
    b._private["_pair_vec4"] = vec4
    out = tb(b)["edge_attr"]
    (gp,) = torch.autograd.grad(out, pos, grad_outputs=torch.from_numpy(g["go"]).to(device))
    pos_c = gd["pos"].clone().requires_grad_(True)
    vec_o, dist_o, cos_o = O.pair_geometry(pos_c, gd["lattice"], gd["batch"], gd["edge_index"],
                                           gd["edge_cell_shift"], gd["triplet_edge_index"])
    sd = {"tb." + k: v for k, v in state_dict_of(g).items()}
    out_o, _ = O.three_body(sd, "tb", O.HyperParams(), torch.from_numpy(g["x"]), torch.from_numpy(g["e"]), dist_o,
                            cos_o, gd["edge_index"], gd["triplet_edge_index"], torch.from_numpy(g["factors"]))
    (gp_o,) = torch.autograd.grad(out_o, pos_c, grad_outputs=torch.from_numpy(g["go"]))
    report("tb.asym.out", out, out_o, 2e-6, 2e-6)
    report("tb.asym.g_pos", gp, gp_o, 5e-5, 5e-5)


def test_conv_operator(device):
    from torch_m3gnet_b200.data.material_graph import get_plan
    from torch_m3gnet_b200.nn.conv import M3GNetConv

    g = golden("conv_op")
    tg = golden("threebody_op")
    gd = graph_dict(tg)
    b = to_batch(gd, device)
    get_plan(b)
    cv = M3GNetConv(3, 64, 64, device=device)
    cv.load_state_dict(state_dict_of(g))
    x = torch.from_numpy(g["x"]).to(device).requires_grad_(True)
    e = torch.from_numpy(g["e"]).to(device).requires_grad_(True)
    h = torch.from_numpy(g["h"]).to(device).requires_grad_(True)
    b["x"], b["edge_attr"], b["edge_weights"] = x, e, h
    out = cv(b)
    report("conv.x_out", out["x"], g["x_out"], 5e-6, 5e-6)
    report("conv.e_out", out["edge_attr"], g["e_out"], 5e-6, 5e-6)
    gx, ge, gh = torch.autograd.grad([out["x"], out["edge_attr"]], [x, e, h],
                                     grad_outputs=[torch.from_numpy(g["gox"]).to(device),
                                                   torch.from_numpy(g["goe"]).to(device)])
    report("conv.gx", gx, g["gx"], 2e-5, 2e-5)
    report("conv.ge", ge, g["ge"], 2e-5, 2e-5)
    report("conv.gh", gh, g["gh"], 2e-5, 2e-5)


@pytest.mark.parametrize("F", [17, 64, 96])
def test_conv_and_readout_generic_width(device, F):
    """Width-agnostic kernels (the reference's own test model uses embedding_dim=17) vs the oracle."""
    from torch_m3gnet_b200.data.material_graph import get_plan
    from torch_m3gnet_b200.nn.conv import M3GNetConv
    from torch_m3gnet_b200.nn.readout import AtomWiseReadout

    gd = graph_dict(golden("small_batch"))
    b = to_batch(gd, device)
    plan = get_plan(b)
    N, E = plan.N, plan.E
    torch.manual_seed(F)
    cv = M3GNetConv(3, F, F)
    ro = AtomWiseReadout(F, 3, scale=2.0)
    sd_c = {"cv." + k: v.detach() for k, v in cv.state_dict().items()}
    sd_r = {"ro." + k: v.detach() for k, v in ro.state_dict().items()}
    x = (0.5 * torch.randn(N, F)).requires_grad_(True)
    e = (0.5 * torch.randn(E, F)).requires_grad_(True)
    h = (0.3 * torch.randn(E, 3)).requires_grad_(True)
    x2, e2 = O.conv(sd_c, "cv", x, e, h, gd["edge_index"])
    elemental = torch.randn(N)
    atomic, stot, tot = O.readout(sd_r, "ro", x2, elemental, gd["batch"], 2, 2.0)
    loss = tot.sum() + (e2 * torch.cos(torch.arange(E * F).reshape(E, F) * 0.37)).sum()
    gx_o, ge_o, gh_o = torch.autograd.grad(loss, [x, e, h])
    cvg, rog = cv.to(device), ro.to(device)
    xg = x.detach().to(device).requires_grad_(True)
    eg = e.detach().to(device).requires_grad_(True)
    hg = h.detach().to(device).requires_grad_(True)
    b["x"], b["edge_attr"], b["edge_weights"], b["elemental_energies"] = xg, eg, hg, elemental.to(device)
    out = rog(cvg(b))
    report(f"F={F} conv.x", out["x"], x2, 5e-6, 5e-6)
    report(f"F={F} conv.e", out["edge_attr"], e2, 5e-6, 5e-6)
    report(f"F={F} readout.atomic", out["scaled_atomic_energies"], atomic, 5e-6, 5e-6)
    report(f"F={F} readout.total", out["total_energy"], tot, 2e-5, 5e-6)
    wgt = torch.cos(torch.arange(E * F).reshape(E, F) * 0.37).to(device)
    lg = out["total_energy"].sum() + (out["edge_attr"] * wgt).sum()
    gx, ge, gh = torch.autograd.grad(lg, [xg, eg, hg])
    report(f"F={F} gx", gx, gx_o, 2e-5, 2e-5)
    report(f"F={F} ge", ge, ge_o, 2e-5, 2e-5)
    report(f"F={F} gh", gh, gh_o, 2e-5, 2e-5)


def test_graph_prep_kernels(device):
    """Integer kernels: scan, CSR builders, row sort, symmetry — bit-exact against numpy."""
    from torch_m3gnet_b200 import _lib

    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 2048, 2049, 100_003):
        v = rng.integers(0, 50, size=n).astype(np.int32)
        t = torch.from_numpy(v).to(device)
        out = torch.empty(n + 1, dtype=torch.int32, device=device)
        work = torch.empty(_lib.scan_work_elems(n), dtype=torch.int32, device=device)
        _lib.call("exclusive_scan_i32", t, out, n, work)
        want = np.concatenate([[0], np.cumsum(v)]).astype(np.int32)
        assert np.array_equal(out.cpu().numpy(), want), f"scan n={n}"
    n, rows = 50_000, 3_001
    keys = rng.integers(0, rows, size=n).astype(np.int32)
    t = torch.from_numpy(keys).to(device)
    ptr = torch.empty(rows + 1, dtype=torch.int32, device=device)
    perm = torch.empty(n, dtype=torch.int32, device=device)
    work = torch.empty(rows + 1 + _lib.scan_work_elems(rows), dtype=torch.int32, device=device)
    _lib.call("csr_by_key", t, n, rows, ptr, perm, work)
    want_perm = np.argsort(keys, kind="stable").astype(np.int32)
    want_ptr = np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=rows))]).astype(np.int32)
    assert np.array_equal(ptr.cpu().numpy(), want_ptr)
    assert np.array_equal(perm.cpu().numpy(), want_perm)
    skeys = torch.from_numpy(np.sort(keys)).to(device)
    ptr2 = torch.empty(rows + 1, dtype=torch.int32, device=device)
    _lib.call("csr_from_sorted", skeys, n, rows, ptr2)
    assert np.array_equal(ptr2.cpu().numpy(), want_ptr)
    flags = torch.empty(4, dtype=torch.int32, device=device)
    _lib.call("check_sorted", skeys, n, rows, flags)
    assert flags[:2].tolist() == [1, 1]
    _lib.call("check_sorted", t, n, rows, flags)
    assert flags[:2].tolist() == [0, 1]
    _lib.call("check_sorted", t, n, rows - 1000, flags)
    assert flags[1].item() == 0


def test_fused_entry_points_equal_their_two_call_forms(device):
    """The entry points the whole-step executor uses to fuse across operator boundaries give exactly what the two
    separate calls give: readout forward + adjoint in one launch, and the first block's three-body edge update that forms
    the EdgeAdjustor's output in-kernel (nn/featurizer.py:84-96 + nn/interaction.py:219-223)."""
    from torch_m3gnet_b200 import _lib
    from torch_m3gnet_b200.nn._functions import sm_count
    from torch_m3gnet_b200.nn.readout import AtomWiseReadout

    torch.manual_seed(11)
    N, B, F = 1000, 7, 64
    f32 = dict(dtype=torch.float32, device=device)
    ro = AtomWiseReadout(F, 3, scale=1.7).to(device)
    w = ro._packed.get()
    x = 0.5 * torch.randn(N, F, **f32)
    elemental = torch.randn(N, **f32)
    batch = torch.sort(torch.randint(0, B, (N,), device=device)).values.to(torch.int32)
    g_total = torch.rand(B, **f32) + 0.5
    names = ("W0dT", "b0d", "W1dT", "b1d", "w2d", "b2d", "W0gT", "b0g", "W1gT", "b1g", "w2g", "b2g")
    back = ("W0d", "W1d", "W0g", "W1g")
    atomic_a, atomic_b = torch.empty(N, **f32), torch.empty(N, **f32)
    gx_a, gx_b = torch.empty(N, F, **f32), torch.empty(N, F, **f32)
    _lib.call("readout_fwd", x, *[w[k] for k in names], elemental, 1.7, N, F, atomic_a)
    _lib.call("readout_bwd", x, *[w[k] for k in names], *[w[k] for k in back], None, None, g_total, batch, 1.7, N, F, gx_a)
    _lib.call("readout_fwd_bwd", x, *[w[k] for k in names], *[w[k] for k in back], elemental, g_total, batch, 1.7, N, F,
              atomic_b, gx_b)
    assert torch.equal(atomic_a, atomic_b) and torch.equal(gx_a, gx_b)

    # e_out = SiLU(h Wa^T) + gated(red): m3g_edge_adjust_fwd + m3g_tb_edge_update  ==  m3g_tb_edge_update_h
    E = 4099  # not a multiple of 32: the last chunk is ragged
    h = 0.3 * torch.randn(E, 3, **f32)
    WaT = 0.4 * torch.randn(3, F, **f32)
    WdT, WgT = 0.5 * torch.randn(9, F, **f32), 0.5 * torch.randn(9, F, **f32)
    red = torch.randn(E, 9, **f32)
    member = torch.rand(E, device=device) < 0.4
    tri_ptr = torch.zeros(E + 1, dtype=torch.int32, device=device)
    tri_ptr[1:] = torch.cumsum(member.to(torch.int32) * 3, 0)
    e0 = torch.empty(E, F, **f32)
    out_a, out_b = torch.empty(E, F, **f32), torch.empty(E, F, **f32)
    _lib.call("edge_adjust_fwd", h, WaT, E, 3, F, e0)
    _lib.call("tb_edge_update", red, tri_ptr, WdT, WgT, e0, E, sm_count(device), out_a)
    _lib.call("tb_edge_update_h", red, tri_ptr, WdT, WgT, h, WaT, E, sm_count(device), out_b)
    assert torch.equal(out_a, out_b)
    # and the gated term itself against torch (member rows only)
    z_d, z_g = red @ WdT, red @ WgT
    want = e0 + torch.where(member[:, None], torch.nn.functional.silu(z_d) * torch.sigmoid(z_g), torch.zeros_like(z_d))
    report("tb_edge_update vs torch", out_a, want, 2e-6, 2e-6)


def test_threebody_mlp_adjoint_over_member_list(device):
    """m3g_tb_mlp_adj (a lane per row over the packed member list, q may overwrite red) against torch autograd of the
    gated 9 -> 64 MLP (nn/interaction.py:219-220, nn/core.py:61-62); rows outside the list are left untouched."""
    from torch_m3gnet_b200 import _lib
    from torch_m3gnet_b200.nn._functions import sm_count

    torch.manual_seed(12)
    f32 = dict(dtype=torch.float32, device=device)
    for E, frac in ((4099, 0.4), (257, 1.0), (40, 0.5)):
        F = 64
        WdT, WgT = 0.5 * torch.randn(9, F, **f32), 0.5 * torch.randn(9, F, **f32)
        red = torch.randn(E, 9, **f32)
        g_e = torch.randn(E, F, **f32)
        members = torch.nonzero(torch.rand(E, device=device) < frac).flatten().to(torch.int32)
        r = red.clone().requires_grad_(True)
        out = torch.nn.functional.silu(r @ WdT) * torch.sigmoid(r @ WgT)
        (want,) = torch.autograd.grad(out, r, grad_outputs=g_e)
        q = torch.full((E, 9), 7.0, **f32)
        _lib.call("tb_mlp_adj", red, g_e, members, int(members.numel()), WdT, WgT, sm_count(device), q)
        idx = members.long()
        report(f"tb_mlp_adj E={E}", q[idx], want[idx], 2e-5, 5e-6)
        rest = torch.ones(E, dtype=torch.bool, device=device)
        rest[idx] = False
        assert torch.all(q[rest] == 7.0)
        alias = red.clone()
        _lib.call("tb_mlp_adj", alias, g_e, members, int(members.numel()), WdT, WgT, sm_count(device), alias)
        assert torch.equal(alias[idx], q[idx]) and torch.equal(alias[rest], red[rest])
