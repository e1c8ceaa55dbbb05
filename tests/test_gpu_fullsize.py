"""GPU: BASELINE.json's full-size configurations through size-independent properties (the oracle cannot run them in
seconds): additivity over structures, structure-order invariance, translation / rotation invariance, Newton's third
law per structure, oracle spot checks on members of the full batch, cell-list == plain sweep on the 32 000-atom cell,
and per-atom three-body kernels == generic CSR kernels on the 1e8-triplet structure."""
import numpy as np
import pytest
import torch

from oracle import m3gnet_oracle as O
from tests.util import golden, report, state_dict_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_x3(device):
    from torch_m3gnet_b200 import build_model

    sd = {k: (v * 3 if k.endswith("weight") else v) for k, v in state_dict_of(golden("c1_default")).items()}
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    model.load_state_dict(sd)   # amplified weights: forces O(0.1) eV/A make the absolute tolerances meaningful
    return model, sd


def _batch(lat, cart, z, sizes, device, **kw):
    from torch_m3gnet_b200 import Batch

    return Batch.from_arrays(lat, cart, z, sizes, 5.0, 4.0, device=device, **kw)


def test_c2_full_batch_properties(device, model_x3):
    """configs[1]: 256 x 108-atom Cu cells = 27 648 atoms, ~1.19 M bonds, ~8.5 M triplets in one call."""
    from torch_m3gnet_b200 import synthetic

    model, sd = model_x3
    lat, cart, z, sizes = synthetic.config2_batch(256)
    b = _batch(lat, cart, z, sizes, device)
    E, T = b["edge_index"].shape[1], b["triplet_edge_index"].shape[1]
    assert b["pos"].shape[0] == 27648 and 1.1e6 < E < 1.3e6 and 8e6 < T < 9e6
    # integer structure: every bond has its reverse, triplet count = sum n3(n3-1)
    ei, sh = b["edge_index"], b["edge_cell_shift"].long()
    key = lambda s, d, c: ((s * 27648 + d) * 7 + (c[:, 0] + 3)) * 49 + (c[:, 1] + 3) * 7 + (c[:, 2] + 3)  # noqa: E731
    fwd, rev = key(ei[0], ei[1], sh), key(ei[1], ei[0], -sh)
    assert torch.equal(torch.sort(fwd).values, torch.sort(rev).values)
    assert int(b["num_triplet_i"].sum()) == T == int(b["num_triplet_ij"].sum())
    out = model(b)
    en, f = out["total_energy"].clone(), out["forces"].clone()
    assert torch.isfinite(en).all() and torch.isfinite(f).all() and f.abs().max() > 1e-2
    # Newton's third law per structure
    fs = f.view(256, 108, 3).sum(1).abs().max().item()
    print(f"[full C2] E={E} T={T} max|F|={f.abs().max():.3e} max|sum F per structure|={fs:.2e}")
    assert fs < 2e-5
    # additivity: members evaluated alone give the batch's numbers.  Every bond row is independent of its neighbours
    # in the tile; the per-atom message sum is grouped by the 32-row blocks of the batch's bond list (conv_tc_fwd mode
    # 2), so a structure's position in the batch moves the grouping: equal to fp32 rounding, and bit for bit with
    # the row-by-row sum (M3G_CONV_MSG_REDUCE=0), checked below
    fmax = f.abs().max().item()

    def same(o_e, ref_e, o_f, ref_f, exact):
        if exact:
            return torch.equal(o_e, ref_e) and torch.equal(o_f, ref_f)
        return ((o_e - ref_e).abs().max().item() <= 2e-6 * ref_e.abs().max().item()
                and (o_f - ref_f).abs().max().item() <= 2e-6 * fmax)

    from torch_m3gnet_b200.nn import conv as conv_mod

    order = np.arange(255, -1, -1)
    cart_r = cart.reshape(256, 108, 3)[order].reshape(-1, 3)
    for exact in (False, True):
        conv_mod.MSG_REDUCE = not exact
        try:
            if exact:
                out = model(b)
                en_x, f_x = out["total_energy"].clone(), out["forces"].clone()
            else:
                en_x, f_x = en, f
            for s in (0, 100, 255):
                a0 = 108 * s
                o1 = model(_batch(lat[s:s + 1], cart[a0:a0 + 108], z[a0:a0 + 108], [108], device))
                assert same(o1["total_energy"], en_x[s:s + 1], o1["forces"], f_x[a0:a0 + 108], exact), (s, exact)
            # structure order: reversed batch = reversed results
            o2 = model(_batch(lat[order], cart_r, z, sizes, device))
            assert same(o2["total_energy"].flip(0), en_x, o2["forces"].view(256, 108, 3).flip(0).reshape(-1, 3), f_x,
                        exact), exact
        finally:
            conv_mod.MSG_REDUCE = True
    dmode = (en_x - en).abs().max().item() / 108
    print(f"[full C2] in-kernel message reduction vs row-by-row sum: |dE|/atom={dmode:.2e} "
          f"max|dF|={(f_x - f).abs().max().item():.2e}")
    assert dmode <= 1e-6
    # oracle spot checks on two members of the full batch (north_star tolerances)
    for s in (7, 200):
        a0 = 108 * s
        g = O.build_graph(lat[s], cart[a0:a0 + 108], z[a0:a0 + 108], 5.0, 4.0)
        ref = O.forward(sd, O.HyperParams(), O.collate([g]), create_graph=False)
        dE = abs(float(en[s].cpu()) - float(ref["total_energy"][0].detach())) / 108
        dF = (f[a0:a0 + 108].cpu() - ref["forces"]).abs().max().item()
        print(f"[full C2] structure {s}: |dE|/atom={dE:.2e} max|dF|={dF:.2e} max|F|={ref['forces'].abs().max():.2e}")
        assert dE <= 1e-5 and dF <= 1e-4


def test_c2_translation_and_rotation_invariance(device, model_x3):
    from torch_m3gnet_b200 import synthetic

    model, _ = model_x3
    lat, cart, z, sizes = synthetic.config2_batch(256)
    out = model(_batch(lat, cart, z, sizes, device))
    en, f = out["total_energy"].clone(), out["forces"].clone()
    # rigid translation out of the home cell (images are relative to the unwrapped coordinates)
    o_t = model(_batch(lat, cart + np.array([13.7, -4.2, 0.9]), z, sizes, device))
    dE = (o_t["total_energy"] - en).abs().max().item() / 108
    dF = (o_t["forces"] - f).abs().max().item()
    print(f"[full C2] translation: |dE|/atom={dE:.2e} max|dF|={dF:.2e}")
    assert dE <= 1e-5 and dF <= 1e-4
    # proper rotation of cell and coordinates: energies equal, forces co-rotate
    q, _ = np.linalg.qr(np.random.default_rng(3).normal(size=(3, 3)))
    q *= np.sign(np.linalg.det(q))
    o_r = model(_batch(lat @ q, cart @ q, z, sizes, device))
    dE = (o_r["total_energy"] - en).abs().max().item() / 108
    dF = (o_r["forces"].double() - f.double() @ torch.as_tensor(q, device=device)).abs().max().item()
    print(f"[full C2] rotation: |dE|/atom={dE:.2e} max|dF|={dF:.2e}")
    assert dE <= 1e-5 and dF <= 1e-4


def test_c4_cell_list_equals_plain_sweep_32000_atoms(device):
    """configs[3]: one 32 000-atom Cu supercell; the binned candidate search must give the plain sweep's bonds."""
    from torch_m3gnet_b200 import Batch, _lib, synthetic

    lat, cart, z = synthetic.fcc_cu_supercell(20, 0.1, 4)
    n = len(cart)
    b = _batch(lat[None], cart, z, [n], device, want_triplet_index=False)
    E = b["edge_index"].shape[1]
    assert n == 32000 and 1.3e6 < E < 1.5e6
    lat64 = torch.as_tensor(lat.reshape(1, 3, 3)).to(device)
    cart64 = torch.as_tensor(cart).to(device)
    atom_ptr = torch.tensor([0, n], dtype=torch.int32, device=device)
    counts = torch.empty(n, dtype=torch.int32, device=device)
    _lib.call("nbr_count", lat64, cart64, atom_ptr, 1, n, 5.0, None, None, None, None, counts)
    edge_ptr = torch.empty(n + 1, dtype=torch.int32, device=device)
    work = torch.empty(_lib.scan_work_elems(n), dtype=torch.int32, device=device)
    _lib.call("exclusive_scan_i32", counts, edge_ptr, n, work)
    assert int(edge_ptr[-1]) == E
    ei = torch.empty((2, E), dtype=torch.int64, device=device)
    sh = torch.empty((E, 3), dtype=torch.int32, device=device)
    dist = torch.empty(E, dtype=torch.float32, device=device)
    member = torch.empty(E, dtype=torch.int32, device=device)
    _lib.call("nbr_fill", lat64, cart64, atom_ptr, 1, n, 5.0, 4.0, None, None, None, None, edge_ptr, E, ei, sh, dist,
              member)
    assert torch.equal(ei, b["edge_index"]) and torch.equal(sh, b["edge_cell_shift"])
    assert torch.equal(dist, b._private["edge_distances_build"])
    n3 = torch.zeros(n, dtype=torch.int64, device=device).index_add_(0, ei[0], member.long())
    assert torch.equal(n3 * (n3 - 1), b["num_triplet_i"])
    assert isinstance(b, Batch)


def test_c5_three_body_atom_path_equals_generic_path_1e8_triplets(device):
    """configs[4]: ~1e8 triplets; the O(n3) moment kernels and the per-atom pair-matrix kernels against the generic CSR
    kernels (different code, different accumulation structure) on the full structure, forward and backward, plus an
    oracle spot check of the same op on a <= 300-atom sub-box at r3 = 5 A with O(1) factors."""
    from torch_m3gnet_b200 import Batch, synthetic
    from torch_m3gnet_b200.nn import interaction
    from torch_m3gnet_b200.nn.invariant import PAIR_VEC4
    from torch_m3gnet_b200.nn._functions import GeometryFn

    lat, cart, z = synthetic.fcc_cu_supercell(23, 0.4, 5)
    b = Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 5.0, device=device, want_triplet_index=False)
    plan = b._plan
    assert plan.T > 9e7 and plan.tri_dense and plan.tri_moment
    tb = interaction.ThreeBodyInteration(5.0, 5.0, 3, 3, 64, 64, device=device)
    fac = torch.rand(3, 3, generator=torch.Generator().manual_seed(3)) + 0.5
    tb.nsb.factors = fac.to(device)
    vec4, dist, _ = GeometryFn.apply(b["pos"], b["lattice"], plan, None)
    vec4 = vec4.detach()
    torch.manual_seed(5)
    x0 = 0.1 * torch.randn(plan.N, 64, device=device)
    e0 = 0.05 * torch.randn(plan.E, 64, device=device)
    go = torch.randn(plan.E, 64, device=device)
    res = {}
    saved = interaction.TB_PATH
    try:
        for path in ("moment", "atom", "generic"):
            interaction.TB_PATH = path
            x, e, v4 = x0.clone().requires_grad_(True), e0.clone().requires_grad_(True), vec4.clone().requires_grad_(True)
            b["x"], b["edge_attr"] = x, e
            b._private.clear()
            b._private[PAIR_VEC4] = v4
            out = tb(b)["edge_attr"]
            gx, ge, gv = torch.autograd.grad(out, [x, e, v4], grad_outputs=go)
            gvec = gv[:, :3] + gv[:, 3:4] * vec4[:, :3] / vec4[:, 3:4]  # total d/dv (the paths split it differently)
            res[path] = [t.detach() for t in (out, gx, ge, gvec)]
            del out, gx, ge, gv
    finally:
        interaction.TB_PATH = saved
        b._private.clear()
    bad = []
    for path in ("moment", "atom"):
        for name, a, g in zip(("out", "g_x", "g_e", "g_vec"), res[path], res["generic"]):
            scale = g.abs().max().item()
            d = (a - g).abs().max().item()
            print(f"[full C5] {path} {name}: max|ref|={scale:.3e} max|diff|={d:.3e}")
            if d > 3e-5 * scale + 1e-7:
                bad.append((path, name, d, scale))
    assert not bad, bad


def test_c5_sub_box_three_body_vs_oracle(device):
    """Oracle spot check at the C5 density: a 4^3-cell sub-box (256 atoms, r_c = r3 = 5 A, jitter 0.4 A: ~45 member
    bonds per atom, ~5e5 triplets) through the moment kernels against the oracle's explicit triplet sum + autograd,
    O(1) factor table, non-unit upstream gradients."""
    from torch_m3gnet_b200 import Batch, synthetic
    from torch_m3gnet_b200.nn import interaction
    from torch_m3gnet_b200.nn.invariant import PAIR_VEC4
    from torch_m3gnet_b200.nn._functions import GeometryFn
    from oracle import m3gnet_oracle as O

    lat, cart, z = synthetic.fcc_cu_supercell(4, 0.4, 5)
    b = Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 5.0, device=device)
    plan = b._plan
    assert plan.tri_moment and plan.max_members > 32
    tb = interaction.ThreeBodyInteration(5.0, 5.0, 3, 3, 64, 64, device=device)
    fac = torch.rand(3, 3, generator=torch.Generator().manual_seed(3)) + 0.5
    tb.nsb.factors = fac.to(device)
    pos = b["pos"].clone().requires_grad_(True)
    vec4, dist, cos = GeometryFn.apply(pos, b["lattice"], plan, b["triplet_edge_index"])
    torch.manual_seed(5)
    x0 = 0.1 * torch.randn(plan.N, 64)
    e0 = 0.05 * torch.randn(plan.E, 64)
    go = torch.randn(plan.E, 64)
    x = x0.to(device).requires_grad_(True)
    b["x"], b["edge_attr"] = x, e0.to(device)
    b._private[PAIR_VEC4] = vec4
    out = tb(b)["edge_attr"]
    gx, gp = torch.autograd.grad(out, [x, pos], grad_outputs=go.to(device))
    # oracle
    hp = O.HyperParams(threebody_cutoff=5.0)
    sd = {"tb." + k: v.detach().cpu() for k, v in tb.state_dict().items()}
    pos_c = b["pos"].cpu().clone().requires_grad_(True)
    x_c = x0.clone().requires_grad_(True)
    ei, tri = b["edge_index"].cpu(), b["triplet_edge_index"].cpu()
    _, dist_o, cos_o = O.pair_geometry(pos_c, b["lattice"].cpu(), b["batch"].cpu(), ei, b["edge_cell_shift"].cpu(), tri)
    out_o, _ = O.three_body(sd, "tb", hp, x_c, e0, dist_o, cos_o, ei, tri, fac)
    gx_o, gp_o = torch.autograd.grad(out_o, [x_c, pos_c], grad_outputs=go)
    report("c5sub.out", out, out_o, 2e-6, 2e-6)
    report("c5sub.g_x", gx, gx_o, 1e-5, 3e-5)
    report("c5sub.g_pos", gp, gp_o, 5e-5, 5e-5)


def _c3_structures():
    """The 1 024 MPF-like structures of BASELINE configs[2] (host worker pool; `spawn`: this process holds a CUDA
    context)."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    from torch_m3gnet_b200 import synthetic

    with ProcessPoolExecutor(max_workers=8, mp_context=mp.get_context("spawn")) as ex:
        return list(ex.map(synthetic.mpf_like_structure, range(1024), chunksize=16))


def test_c3_full_ragged_batch_1024_structures(device, model_x3):
    """configs[2] at full size: 1 024 ragged multi-element structures (~1.1e5 atoms, ~3.8e6 bonds, ~3.2e7 triplets) in
    one call: index identities, Newton's third law per structure, additivity (members evaluated alone), structure-order
    invariance, and oracle spot checks on the largest structure, the smallest, the one with the most species and the
    densest one."""
    from torch_m3gnet_b200 import Batch

    model, sd = model_x3
    structs = _c3_structures()
    sizes = [len(s[1]) for s in structs]
    lat = np.stack([s[0] for s in structs])
    cart = np.concatenate([s[1] for s in structs])
    z = np.concatenate([s[2] for s in structs])
    off = np.concatenate([[0], np.cumsum(sizes)])
    b = Batch.from_arrays(lat, cart, z, sizes, 5.0, 4.0, device=device)
    N, E, T = b["pos"].shape[0], b["edge_index"].shape[1], b["triplet_edge_index"].shape[1]
    assert len(structs) == 1024 and 1.0e5 < N < 1.3e5 and E > 3e6 and T > 2e7
    assert int(b["num_triplet_i"].sum()) == T == int(b["num_triplet_ij"].sum())
    # every bond has its reverse (same structure, opposite image)
    ei, sh = b["edge_index"], b["edge_cell_shift"].long()
    key = lambda s, d, c: ((s * N + d) * 9 + (c[:, 0] + 4)) * 81 + (c[:, 1] + 4) * 9 + (c[:, 2] + 4)  # noqa: E731
    assert int(sh.abs().max()) <= 4
    assert torch.equal(torch.sort(key(ei[0], ei[1], sh)).values, torch.sort(key(ei[1], ei[0], -sh)).values)
    out = model(b)
    en, f = out["total_energy"].clone(), out["forces"].clone()
    assert torch.isfinite(en).all() and torch.isfinite(f).all()
    fmax = f.abs().max().item()
    batch_idx = b["batch"]
    fsum = torch.zeros((1024, 3), device=device).index_add_(0, batch_idx, f).abs().max().item()
    print(f"[full C3] N={N} E={E} T={T} max|F|={fmax:.3e} max|sum F per structure|={fsum:.2e}")
    assert fsum < 1e-4 * max(fmax, 1.0)
    n_species = [len(set(s[2].tolist())) for s in structs]
    density = [len(s[1]) / abs(np.linalg.det(s[0])) for s in structs]
    picks = sorted({int(np.argmax(sizes)), int(np.argmin(sizes)), int(np.argmax(n_species)), int(np.argmax(density)),
                    511})
    for s in picks:
        a0, a1 = off[s], off[s + 1]
        o1 = model(Batch.from_arrays(lat[s:s + 1], cart[a0:a1], z[a0:a1], [sizes[s]], 5.0, 4.0, device=device))
        dE = (o1["total_energy"][0] - en[s]).abs().item() / sizes[s]
        dF = (o1["forces"] - f[a0:a1]).abs().max().item()
        g = O.build_graph(lat[s], cart[a0:a1], z[a0:a1], 5.0, 4.0)
        ref = O.forward(sd, O.HyperParams(), O.collate([g]), create_graph=False)
        dEo = abs(float(en[s].cpu()) - float(ref["total_energy"][0].detach())) / sizes[s]
        dFo = (f[a0:a1].cpu() - ref["forces"]).abs().max().item()
        fo = ref["forces"].abs().max().item()
        print(f"[full C3] structure {s} ({sizes[s]} atoms, {n_species[s]} species): alone vs batch |dE|/atom={dE:.1e} "
              f"|dF|={dF:.1e}; oracle |dE|/atom={dEo:.2e} max|dF|={dFo:.2e} max|F|={fo:.2e}")
        assert dE <= 1e-6 and dF <= 2e-6 * max(fmax, 1.0)
        assert dEo <= 1e-5 and dFo <= 1e-4 + 1e-3 * fo
    # structure order: reversed batch = reversed results (to fp32 rounding: see test_c2_full_batch_properties)
    order = np.arange(1023, -1, -1)
    cart_r = np.concatenate([structs[s][1] for s in order])
    z_r = np.concatenate([structs[s][2] for s in order])
    o2 = model(Batch.from_arrays(lat[order], cart_r, z_r, [sizes[s] for s in order], 5.0, 4.0, device=device))
    assert (o2["total_energy"].flip(0) - en).abs().max().item() <= 2e-6 * en.abs().max().item()
    f_back = torch.cat([o2["forces"][int(a):int(bb)] for a, bb in
                        zip(np.concatenate([[0], np.cumsum([sizes[s] for s in order])])[:-1][::-1],
                            np.cumsum([sizes[s] for s in order])[::-1])])
    assert (f_back - f).abs().max().item() <= 2e-6 * max(fmax, 1.0)
