#!/usr/bin/env python
"""bench.py — energy+forces atom-steps/s of the M3GNet hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N=1): BASELINE.json configs[1] — 256 randomly perturbed 108-atom FCC Cu supercells, default M3GNet
(3 blocks, 64 units, r_c=5, r3=4), random-init weights, fp32.  One "step" = one ``model(batch)`` call = energies
+ forces (+ virial) of all 27 648 atoms.  For N>1 every rank processes its own 256-structure batch
(structures are independent: sharded by graph, no data-path collective; "scaling": "weak").

  value  atom-steps/s with the batch resident in HBM, CUDA-event timed, max over ranks
  e2e    the same call from pinned HOST buffers: H2D of the whole graph + model(batch) + D2H of energies/forces
  roofline / rooflines   per-kernel CUDA-event durations (separate instrumented pass) against the measured peaks
  cpu_baseline           the oracle port of the reference's CPU path on the host cores (bounded sample)

``--impl reference`` times the reference's CPU implementation (oracle port, all host threads) on a bounded
sample of the same workload and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HP = dict(cutoff=5.0, threebody_cutoff=4.0, l_max=3, n_max=3, num_types=95, embedding_dim=64, num_blocks=3)
N_STRUCT = 256
WORKLOAD = "C2: 256 x 108-atom FCC Cu (3x3x3 cells, a=3.615, jitter +-0.1 A), default M3GNet, energy+forces"
METRIC = "energy+forces atom-steps/sec"
UNIT = "atom-steps/s"
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4 (not in MEASURED_PEAKS.json; nominal at max clock)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ----------------------------------------------------------------------------------------------------------
def build_inputs(device, seed0: int):
    """Graph build on the GPU (one-time, outside every timed region) + pinned host copies for the e2e leg."""
    import torch_m3gnet_b200 as m3g
    from torch_m3gnet_b200 import synthetic

    lat, cart, z, sizes = synthetic.config2_batch(N_STRUCT, first_seed=seed0)
    t0 = time.time()
    batch = m3g.Batch.from_arrays(lat, cart, z, sizes, HP["cutoff"], HP["threebody_cutoff"], device=device)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    host = {}
    for k in ("pos", "atom_types", "num_triplet_i", "edge_index", "edge_cell_shift", "num_triplet_ij",
              "triplet_edge_index", "lattice", "batch"):
        host[k] = batch[k].cpu().pin_memory()
    return batch, host, build_s


def batch_from_host(host, device):
    import torch_m3gnet_b200 as m3g

    dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
    b = m3g.Batch(pos=dev["pos"], atom_types=dev["atom_types"], num_triplet_i=dev["num_triplet_i"],
                  edge_index=dev["edge_index"], edge_cell_shift=dev["edge_cell_shift"],
                  num_triplet_ij=dev["num_triplet_ij"], triplet_edge_index=dev["triplet_edge_index"],
                  lattice=dev["lattice"])
    b["batch"] = dev["batch"]
    return b


# ABI entry -> kernel name in the committed ncu launch list (profiles/traffic.json, written by tools/summarize_ncu.py)
TRAFFIC_KERNEL = {
    "conv_tc_bwd": "conv_tc_bwd2_kernel", "conv_tc_bwd_saved": "conv_tc_bwds_kernel",
    "conv_tc_fwd": "conv_tc4_fwd_kernel", "tb_atom_fwd": "tb_atom_fwd_kernel",
    "tb_atom_bwd": "tb_atom_bwd_kernel", "conv_gather_gz": "conv_gather_gz128_kernel",
    "segment_sum_add": "segment_sum_add_kernel", "tb_sigma_fwd": "tb_sigma_fwd_kernel",
    "tb_sigma_bwd": "tb_sigma_bwd_kernel", "tb_edge_basis_fwd": "tb_edge_basis_fwd_kernel<3, 3>",
    "tb_edge_basis_bwd": "tb_edge_basis_bwd_kernel<3, 3>",
}


def load_traffic():
    """Measured DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of the same bench command on
    B200; static file, NOT measured in this run) or {} when the file is absent."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return {}
    return json.load(open(path)).get("bytes_per_launch", {})


def kernel_model(E, T, N, F=64, R=3, D=9):
    """Algorithmic bytes / flops per launch (restated in DESIGN.md §4).

    The three-body op is reported as a GROUP (all kernels of its forward, resp. backward, summed): the survey's
    byte formula covers the whole op, not only the triplet-reduction kernel."""
    mlp_mac = F * 2 * F + 2 * F * F  # split first layer (e·W1e: F x 2F) + two F x F second layers, per edge
    single = {
        # name: (bound, algorithmic bytes, algorithmic flops)
        "conv_mlp_fwd": ("tensor", 532 * E // 2 + 512 * N // 2, 2 * E * (mlp_mac + 2 * R * F)),
        "conv_mlp_bwd": ("tensor", 800 * E // 2 + 768 * N // 2, 2 * E * (2 * mlp_mac + F * 2 * F)),
    }
    # tensor-core pair actually used for F = 64: the forward also writes 1 KB/edge of activations, the backward reads
    # them instead of recomputing the forward GEMMs (flops = adjoint GEMMs only: 4 x 64x64 per edge)
    single["conv_tc_fwd"] = ("tensor", 532 * E // 2 + 512 * N // 2 + 1024 * E, 2 * E * (mlp_mac + 2 * R * F))
    single["conv_tc_bwd"] = single["conv_mlp_bwd"]
    single["conv_tc_bwd_saved"] = ("tensor", 800 * E // 2 + 768 * N // 2 + 1024 * E, 2 * E * (4 * F * F))
    groups = {
        "threebody_fwd": (("tb_sigma_fwd", "tb_edge_basis_fwd", "tb_reduce_fwd", "tb_reduce_fwd_fast", "tb_atom_fwd"),
                          "hbm", 4 * T + 536 * E + 36 * N),
        "threebody_bwd": (("tb_gate_bwd", "tb_gate_bwd_fast", "tb_reduce_bwd", "tb_reduce_bwd_sym", "tb_atom_bwd",
                           "tb_edge_basis_bwd", "tb_sigma_bwd"), "hbm", 4 * T + 332 * E + 72 * N),
    }
    return single, groups


def profile_pass(model, batch, steps, peaks):
    """Per-kernel durations with CUDA events on the launching stream (separate pass, not the headline timing)."""
    from torch_m3gnet_b200 import _lib

    _lib.PROFILE = {}
    for _ in range(steps):
        model(batch)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    per = {}
    for name, evs in prof.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        per[name] = dict(calls_per_step=len(ms) / steps, ms_per_step=sum(ms) / steps, avg_ms=sum(ms) / len(ms))
    total = sum(v["ms_per_step"] for v in per.values())
    plan = batch._plan
    km, groups = kernel_model(plan.E, plan.T, plan.N)
    tensor_peak = peaks["bf16_sustained"] or peaks["bf16"]
    traffic = load_traffic()
    rooflines = []
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        entry = dict(kernel="m3g_" + name, share=v["ms_per_step"] / total, avg_ms=v["avg_ms"],
                     launches_per_step=v["calls_per_step"], traffic=traffic.get(TRAFFIC_KERNEL.get(name, "")))
        if name in km:
            bound, nbytes, flops = km[name]
            sec = v["avg_ms"] * 1e-3
            a = flops / sec / 1e12
            gbs = nbytes / sec / 1e9
            if gbs / peaks["hbm"] >= a / tensor_peak:
                # the narrow (F = 64) gated MLPs sit closer to the HBM roof than to the tensor roof
                entry.update(bound="hbm", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                             tensor_tflops=a, tensor_frac=a / tensor_peak, tf32x3_frac=a / (tensor_peak / 2 / 3),
                             note="algorithmic bytes (DESIGN.md 4, incl. the 1 KB/edge of saved activations) against "
                                  "the measured HBM peak; tensor_frac = algorithmic flops against the measured bf16 "
                                  "peak, tf32x3_frac against bf16_peak/2/3 (tcgen05 kind::tf32, 3xTF32 split)")
            else:
                entry.update(bound="tensor", achieved=a, peak=tensor_peak, unit="TFLOP/s", frac=a / tensor_peak,
                             tf32x3_frac=a / (tensor_peak / 2 / 3), hbm_frac=gbs / peaks["hbm"],
                             note="tcgen05 kind::tf32, 3 passes (3xTF32 split); peak = measured sustained bf16 cuBLAS; "
                                  "tf32x3_frac = against bf16_peak/2/3; hbm_frac = algorithmic bytes vs measured HBM")
        rooflines.append(entry)
    for gname, (members, bound, nbytes) in groups.items():
        present = [m for m in members if m in per]
        if not present:
            continue
        # one op instance = one launch of each member kernel
        sec = sum(per[m]["avg_ms"] for m in present) * 1e-3
        a = nbytes / sec / 1e9
        tr = [traffic.get(TRAFFIC_KERNEL.get(m, "")) for m in present]
        rooflines.append(dict(kernel=gname + " (" + "+".join("m3g_" + m for m in present) + ")",
                              traffic=(sum(tr) if all(t is not None for t in tr) else None),
                              share=sum(per[m]["ms_per_step"] for m in present) / total, avg_ms=sec * 1e3,
                              launches_per_step=per[present[0]]["calls_per_step"], bound=bound, achieved=a,
                              peak=peaks["hbm"], unit="GB/s", frac=a / peaks["hbm"],
                              triplets_per_s=plan.T / sec))
    rooflines.sort(key=lambda r: -r["share"])
    return rooflines, total


def cpu_baseline(sd_cpu, n_sample_structs=2, reps=3):
    """The oracle port of the reference's CPU path on a bounded sample (structures of the same workload)."""
    from oracle import m3gnet_oracle as O
    from torch_m3gnet_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = O.HyperParams(**HP)
    graphs = []
    for s in range(n_sample_structs):
        lat, cart, z = synthetic.fcc_cu_supercell(3, 0.1, s)
        graphs.append(O.build_graph(lat, cart, z, HP["cutoff"], HP["threebody_cutoff"]))
    g = O.collate(graphs)
    fac = O.bessel_factors(hp.scaled_cutoff, hp.l_max, hp.n_max)
    n_atoms = g["pos"].shape[0]
    O.forward(sd_cpu, hp, {k: v.clone() for k, v in g.items()}, factors=fac)  # warm-up
    best = float("inf")
    for _ in range(reps):
        gi = {k: v.clone() for k, v in g.items()}
        t0 = time.perf_counter()
        O.forward(sd_cpu, hp, gi, factors=fac)  # create_graph=True as nn/gradient.py:33 does
        best = min(best, time.perf_counter() - t0)
    return dict(value=n_atoms / best, unit=UNIT, cores=cores, kind="port",
                sample=f"{n_sample_structs} of the {N_STRUCT} structures ({n_atoms} atoms) per call, best of {reps}; "
                       f"graph given; oracle/m3gnet_oracle.py (torch CPU fp32, create_graph=True)",
                seconds_per_call=best)


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import m3gnet_oracle as O
    from torch_m3gnet_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = O.HyperParams(**HP)
    torch.manual_seed(0)
    sd = O.init_params(hp, seed=0)
    n_sample = 2
    graphs = []
    for s in range(n_sample):
        lat, cart, z = synthetic.fcc_cu_supercell(3, 0.1, s)
        graphs.append(O.build_graph(lat, cart, z, HP["cutoff"], HP["threebody_cutoff"]))
    g = O.collate(graphs)
    fac = O.bessel_factors(hp.scaled_cutoff, hp.l_max, hp.n_max)
    n_atoms = g["pos"].shape[0]
    for _ in range(args.warmup):
        O.forward(sd, hp, {k: v.clone() for k, v in g.items()}, factors=fac)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.forward(sd, hp, {k: v.clone() for k, v in g.items()}, factors=fac)
    dt = time.perf_counter() - t0
    value = n_atoms * args.steps / dt
    sample = (f"each step = {n_sample} of the {N_STRUCT} structures ({n_atoms} atoms), graph given; "
              f"atom-steps/s is per-atom so the sample scales linearly")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD, sample=sample),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (used under ncu only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if distributed:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)

    import torch_m3gnet_b200 as m3g
    from torch_m3gnet_b200 import _lib

    peaks = load_peaks()
    torch.manual_seed(0)
    model = m3g.build_model(**HP, device=None)  # CPU init (seed 0) so that every rank holds the same weights
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = m3g.build_model(**HP, device=device)
    model.load_state_dict(sd_cpu)

    # rank r works on structures seeded r*256 .. r*256+255 (sharded by graph)
    batch, host, build_s = build_inputs(device, seed0=rank * N_STRUCT)
    plan = batch._plan
    n_atoms, n_edges, n_tri = plan.N, plan.E, plan.T

    def sync_all():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        model(batch)
    sync_all()
    l0 = _lib.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            model(batch)
        ev1.record()
        sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = (_lib.LAUNCHES - l0) // args.steps
    t_ms = torch.tensor([ms], device=device)
    if distributed:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    total_atoms = n_atoms * world
    value = total_atoms * args.steps / (ms_max * 1e-3)

    # ---------------- end to end from pinned host buffers ----------------
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    e_host = torch.empty(N_STRUCT, dtype=torch.float32).pin_memory()
    f_host = torch.empty((n_atoms, 3), dtype=torch.float32).pin_memory()
    d2h = e_host.numel() * 4 + f_host.numel() * 4

    # Two-stage pipeline, as a serving loop would run it: while step k computes on the main stream, the side stream
    # uploads step k+1's graph from pinned host memory and derives its plan (integer kernels).  Every step's H2D copy,
    # plan build, model call and D2H of energies + forces happens inside the timed region; the result of step k is on
    # the host (stream synchronised) before step k+1 is launched.
    from torch_m3gnet_b200.data.material_graph import get_plan

    main_stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=device)

    def stage_in():
        with torch.cuda.stream(side):
            b = batch_from_host(host, device)
            get_plan(b)
            ready = torch.cuda.Event()
            ready.record(side)
        return b, ready

    def e2e_run(n):
        nxt = stage_in() if n > 0 else None
        for k in range(n):
            b, ready = nxt
            main_stream.wait_event(ready)
            out = model(b)
            e_host.copy_(out["total_energy"], non_blocking=True)
            f_host.copy_(out["forces"], non_blocking=True)
            nxt = stage_in() if k + 1 < n else None  # overlaps the kernels of step k
            main_stream.synchronize()               # step k's energies and forces are on the host
            del b, out                               # released only after the main stream has finished with them

    e2e_steps = 0 if args.no_e2e else args.steps
    e2e_run(0 if args.no_e2e else 2)
    sync_all()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    sync_all()
    e2e_s = torch.tensor([max(time.perf_counter() - t0, 1e-9)], device=device)
    if distributed:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = total_atoms * e2e_steps / float(e2e_s.item())

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD, atoms_per_gpu=n_atoms, edges_per_gpu=n_edges,
                            triplets_per_gpu=n_tri, structures_per_gpu=N_STRUCT,
                            parallelism=f"graph-sharded x{world}, no data-path collective",
                            l2="per-step working set (edge features 6 x 297 MB + saved activations) >> 126 MB L2",
                            graph_build_s=build_s),
                triplets_per_s=n_tri * world * args.steps / (ms_max * 1e-3),
                clocks=clocks.summary(),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         note="pinned host graph -> H2D -> plan -> model(batch) -> D2H energies+forces; the upload + plan of step k+1 "
                              "overlap the kernels of step k (side stream), every step's copies are inside the timed region"),
                gpu_launches=int(launches))
    if not args.no_profile:
        rooflines, kernel_ms = profile_pass(model, batch, 2, peaks)
        dominant = next((r for r in rooflines if "bound" in r), None)
        if dominant is not None:
            line["roofline"] = dict(bound=dominant["bound"], achieved=dominant["achieved"], peak=dominant["peak"],
                                    unit=dominant["unit"], frac=dominant["frac"], traffic=dominant.get("traffic"),
                                    kernel=dominant["kernel"], peak_source=peaks["source"] + " (MEASURED_PEAKS.json)",
                                    tensor_frac=dominant.get("tensor_frac"),
                                    note="dominant kernel of the step; algorithmic bytes or flops (DESIGN.md 4) / "
                                         "CUDA-event duration against the measured peak of the nearer roof; traffic = "
                                         "ncu DRAM bytes per launch from profiles/traffic.json (static, same command)")
        line["rooflines"] = rooflines[:14]
        line["kernel_ms_per_step"] = kernel_ms
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(sd_cpu)
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
