"""Isolated ThreeBodyInteraction fwd+bwd microbenchmark (BASELINE.json configs[4] / SURVEY.md §8(d) "C5").

Dense triplet-heavy structure: 23^3 conventional FCC Cu cells (48 668 atoms, 83.1 A box), positions jittered
U(-0.4, 0.4) A (default_rng(5)), r_c = r3 = 5 A  ->  E ~ 2.2e6 bonds, T ~ 1e8 triplets.  Times the three-body op
alone (forward + backward to x, e and the bond vectors) with CUDA events and reports triplets/s and the fraction of the
HBM roofline given by the algorithmic-byte formula of SURVEY.md §8(d):  fwd+bwd = 8 T + 868 E + 108 N bytes.

  python tools/bench_threebody.py [--cells 23] [--steps 10] [--path atom|fast|generic]

Prints one JSON line.  `--cells 8` is a quick smoke size.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch_m3gnet_b200 as m3g  # noqa: E402
from torch_m3gnet_b200 import _lib, synthetic  # noqa: E402
from torch_m3gnet_b200.nn import interaction  # noqa: E402
from torch_m3gnet_b200.nn._functions import ThreeBodyFn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=23)
    ap.add_argument("--jitter", type=float, default=0.4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--path", default="moment", choices=["moment", "atom", "fast", "generic"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    interaction.TB_PATH = args.path
    lat, cart, z = synthetic.fcc_cu_supercell(args.cells, args.jitter, 5)
    t0 = time.time()
    batch = m3g.Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 5.0, device=dev, want_triplet_index=False)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    plan = batch._plan
    N, E, T = plan.N, plan.E, plan.T
    tb = interaction.ThreeBodyInteration(5.0, 5.0, 3, 3, 64, 64, device=dev)
    tb.nsb.factors = (torch.rand(3, 3, generator=torch.Generator().manual_seed(3)) + 0.5).to(dev)  # O(1) table (Q1)
    w = tb._packed.get()
    vec4 = torch.empty((E, 4), device=dev)
    dist = torch.empty(E, device=dev)
    _lib.call("geometry_fwd", batch["pos"], batch["lattice"], plan.batch, plan.src, plan.dst, plan.shift, E, vec4, dist)
    torch.manual_seed(5)
    x = (0.1 * torch.randn(N, 64, device=dev)).requires_grad_(True)
    e = (0.05 * torch.randn(E, 64, device=dev)).requires_grad_(True)
    v4 = vec4.clone().requires_grad_(True)
    go = torch.randn(E, 64, device=dev)

    from torch_m3gnet_b200.nn.invariant import PAIR_VEC4

    def step():
        # the module call: includes the per-step radial tables (m3g_tb_radial) of the moment path
        batch._private.clear()
        batch._private[PAIR_VEC4] = v4
        batch["x"], batch["edge_attr"] = x, e
        out = tb(batch)["edge_attr"]
        torch.autograd.grad(out, [x, e, v4], grad_outputs=go)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    # per-kernel split of one step
    _lib.PROFILE = {}
    step()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    kernels = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in prof.items()}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    alg_bytes = 8 * T + 868 * E + 108 * N
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(
        metric="three-body triplets/sec (isolated ThreeBodyInteraction fwd+bwd)", value=T / (ms * 1e-3),
        unit="triplets/s", ms_per_step=ms, n_gpus=1, steps=args.steps, dtype="f32", data="synthetic",
        config=dict(workload=f"C5: {args.cells}^3 FCC Cu cells, jitter +-{args.jitter} A, r_c = r3 = 5 A",
                    atoms=N, bonds=E, triplets=T, triplets_per_bond=T / max(E, 1), max_members=plan.max_members,
                    path=(args.path if (args.path in ("atom", "moment") and plan.tri_dense) else args.path), graph_build_s=build_s,
                    l2="edge features (E x 256 B read + write) exceed the 126 MB L2"),
        roofline=dict(bound="hbm", achieved=achieved, peak=hbm, unit="GB/s", frac=achieved / hbm, traffic=None,
                      algorithmic_bytes=alg_bytes, formula="8 T + 868 E + 108 N (SURVEY.md 8(d))"),
        kernels_ms=kernels)))


if __name__ == "__main__":
    main()
