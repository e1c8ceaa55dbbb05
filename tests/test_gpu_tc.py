"""GPU: tcgen05 (UMMA) plumbing self-test and the tensor-core gated-MLP path against the generic fp32 kernels."""
import numpy as np
import pytest
import torch

from tests.util import golden, graph_dict, report, state_dict_of, to_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(64, 64), (128, 64), (64, 128)])
def test_umma_selftest(device, rows, cols):
    """One 128-row tile through tcgen05.mma kind::tf32 (K-major SWIZZLE_128B operands, accumulator in TMEM)."""
    from torch_m3gnet_b200 import _lib

    torch.manual_seed(rows * 7 + cols)
    W = torch.randn(rows, cols, device=device)
    A = torch.randn(128, cols, device=device)
    hi = torch.empty(rows * cols, device=device)
    lo = torch.empty(rows * cols, device=device)
    _lib.call("tc_pack_b", W, rows, cols, hi, lo)
    want = (A.double() @ W.double().t()).float()
    for a_tmem in (0, 1):
        for passes, tol in ((3, 1e-5), (1, 3e-3)):
            out = torch.full((128, rows), float("nan"), device=device)
            _lib.call("tc_selftest", A, hi, lo, rows, cols, passes, a_tmem, out)
            torch.cuda.synchronize()
            report(f"umma rows={rows} cols={cols} passes={passes} a_tmem={a_tmem}", out, want, 0, tol)


@pytest.mark.parametrize("path,bwd_variant", [("tc3", 4), ("tc3", 2), ("tc1", 4)])
def test_conv_tc_matches_fma(device, path, bwd_variant):
    """tensor-core gated MLPs (backward from saved activations = 4, with the forward recomputed = 2) against the generic
    fp32 FMA kernels and the live-reference fixture"""
    from torch_m3gnet_b200.data.material_graph import get_plan
    from torch_m3gnet_b200.nn import conv as conv_mod
    from torch_m3gnet_b200.nn.conv import M3GNetConv

    g = golden("conv_op")
    gd = graph_dict(golden("threebody_op"))
    b = to_batch(gd, device)
    get_plan(b)
    cv = M3GNetConv(3, 64, 64, device=device)
    cv.load_state_dict(state_dict_of(g))
    outs = {}
    old, old_bwd = conv_mod.CONV_PATH, conv_mod.TC_BWD_VARIANT
    try:
        conv_mod.TC_BWD_VARIANT = bwd_variant
        for p in ("fma", path):
            conv_mod.CONV_PATH = p
            b["x"] = torch.from_numpy(g["x"]).to(device)
            b["edge_attr"] = torch.from_numpy(g["e"]).to(device)
            b["edge_weights"] = torch.from_numpy(g["h"]).to(device)
            x_in = b["x"].requires_grad_(True)
            e_in = b["edge_attr"].requires_grad_(True)
            h_in = b["edge_weights"].requires_grad_(True)
            out = cv(b)
            grads = torch.autograd.grad([out["x"], out["edge_attr"]], [x_in, e_in, h_in],
                                        grad_outputs=[torch.from_numpy(g["gox"]).to(device),
                                                      torch.from_numpy(g["goe"]).to(device)])
            torch.cuda.synchronize()
            outs[p] = (out["x"].detach().clone(), out["edge_attr"].detach().clone()) + tuple(grads)
    finally:
        conv_mod.CONV_PATH, conv_mod.TC_BWD_VARIANT = old, old_bwd
    tol = 1e-5 if path == "tc3" else 5e-3
    report(f"conv {path} e_out vs fma", outs[path][1], outs["fma"][1], tol, tol)
    report(f"conv {path} x_out vs fma", outs[path][0], outs["fma"][0], tol * 4, tol)
    report(f"conv {path} e_out vs reference", outs[path][1], g["e_out"], tol * 2, tol)
    report(f"conv {path} x_out vs reference", outs[path][0], g["x_out"], tol * 4, tol)
    for i, name in ((2, "gx"), (3, "ge"), (4, "gh")):
        report(f"conv {path} {name} vs fma", outs[path][i], outs["fma"][i], tol * 4, tol * 4)
        report(f"conv {path} {name} vs reference", outs[path][i], g[name], tol * 4, tol * 4)
