"""Trajectory loop benchmark: velocity-Verlet MD steps/s with the Verlet (skin) list against
a full neighbour-list rebuild per frame.

  python tools/bench_md.py [--cells 20] [--steps 50] [--skin 0.5] [--dt 2.0] [--temperature 600]

--cells 20 is BASELINE.json configs[3] (32 000-atom Cu supercell); --cells 2 the 32-atom cell of configs[0].
Prints one JSON line: MD steps/s and atom-steps/s for both arms, graph-update time per frame, rebuild count.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch_m3gnet_b200 as m3g  # noqa: E402
from torch_m3gnet_b200 import synthetic  # noqa: E402


def timed(fn, n):
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=20)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--skin", type=float, default=0.5)
    ap.add_argument("--dt", type=float, default=2.0)
    ap.add_argument("--temperature", type=float, default=600.0)
    ap.add_argument("--jitter", type=float, default=0.05)
    ap.add_argument("--replay", action="store_true", help="Verlet arm with graph_replay=True (cold / small systems)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lat, cart, z = synthetic.fcc_cu_supercell(args.cells, args.jitter, 4)
    n = len(cart)
    torch.manual_seed(0)
    model = m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=dev)
    mass = np.full(n, 63.546)
    sigma = np.sqrt(8.617333e-5 * args.temperature * m3g.calculator.ACC_UNIT / 63.546)  # A/fs per component
    v0 = np.random.default_rng(0).normal(0.0, sigma, size=cart.shape)
    v0 -= v0.mean(axis=0)

    def run_arm(force_rebuild: bool):
        calc = m3g.M3GNetCalculator(model, 5.0, 4.0, skin=args.skin, device=dev, round_allocations=True,
                                    graph_replay=args.replay and not force_rebuild)
        md = m3g.VelocityVerlet(calc, lat, cart, z, mass, dt=args.dt, velocities=v0)

        def one():
            if force_rebuild:
                calc.neighbor_list._ref = None  # candidates rebuilt every frame: the skin list disabled
            md.step(1)

        for _ in range(args.warmup):
            one()
        r0 = calc.neighbor_list.n_rebuilds
        ms = timed(one, args.steps)
        return ms, calc.neighbor_list.n_rebuilds - r0, calc, md

    # Every arm integrates the same trajectory from the same start.  The first pass also warms the caching allocator
    # (each frame has its own bond / triplet count, see calculator.stabilise_allocator); it is reported as "cold".
    ms_cold, _, _, _ = run_arm(False)
    ms_full, _, _, _ = run_arm(True)
    ms_verlet, rebuilds, calc, md = run_arm(False)
    # graph update alone (filter + triplets + plan) on the last frame, and a full rebuild of the same frame
    vl = calc.neighbor_list
    pos = md.pos.clone()
    ms_update = timed(lambda: vl.update(pos), 20)
    pos_h = pos.cpu().numpy()

    def fresh():
        return m3g.Batch.from_arrays(lat[None], pos_h, z, [n], 5.0, 4.0, device=dev, want_triplet_index=False)

    def fresh_dev():  # without the host->device copy of the coordinates: sweep + triplets + plan only
        vl._ref = None
        vl.update(pos)

    fresh()
    ms_fresh = timed(fresh, 20)
    ms_rebuild = timed(fresh_dev, 10)
    b = vl.update(pos)
    print(json.dumps(dict(
        metric="velocity-Verlet MD steps/sec (energy+forces per step)", unit="steps/s", n_gpus=1, steps=args.steps,
        value=1e3 / ms_verlet, atom_steps_per_s=n * 1e3 / ms_verlet, ms_per_step=ms_verlet,
        rebuild_every_frame=dict(value=1e3 / ms_full, ms_per_step=ms_full),
        first_pass_cold_allocator=dict(value=1e3 / ms_cold, ms_per_step=ms_cold),
        graph_ms=dict(verlet_update=ms_update, candidate_rebuild_plus_update=ms_rebuild, from_arrays_host_coords=ms_fresh),
        candidate_rebuilds_in_timed_steps=rebuilds, graph_replays=calc.n_replays, graph_captures=calc.n_captures,
        config=dict(workload=f"{args.cells}^3 FCC Cu cells ({n} atoms), T0={args.temperature} K, dt={args.dt} fs, "
                             f"skin={args.skin} A, jitter={args.jitter} A", atoms=n, bonds=b._plan.E, triplets=b._plan.T,
                    candidates=vl.C), dtype="f32", data="synthetic")))


if __name__ == "__main__":
    main()
