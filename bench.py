#!/usr/bin/env python
"""bench.py — energy+forces atom-steps/s of the M3GNet hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1   headline = BASELINE.json configs[1] ("C2"): 256 randomly perturbed 108-atom FCC Cu supercells, default M3GNet
        (3 blocks, 64 units, r_c = 5, r3 = 4), random-init weights, fp32.  One "step" = one ``model(batch)`` call =
        energies + forces (+ virial) of all 27 648 atoms.  The line also carries the single-GPU figures of the two
        sharded configurations ("c3", "dd") so that the N > 1 lines have their own baselines.
N > 1   headline = configs[2] ("C3"): 1 024 MPF-like ragged multi-element structures sharded BY STRUCTURE over the ranks
        (longest-processing-time assignment on a predicted cost, no data-path collective): STRONG scaling, the total
        work is fixed.  "dd" = configs[3] ("C4"): one 32 000-atom Cu cell, spatially domain-decomposed over the ranks
        with per-block halo exchanges of the node features over NCCL (torch_m3gnet_b200/domain.py).

  value      atom-steps/s with the batch resident in HBM, CUDA-event timed, max over ranks
  e2e        the call a user makes, from pinned HOST buffers: coordinates / cells / atomic numbers -> H2D -> graph build
             on the GPU (neighbour list, triplets) -> model -> D2H of energies and forces, every step
  e2e_graph_given   (N = 1) the same with a ready-made reference-format graph in pinned host memory (all index tensors
             uploaded every step, no graph build)
  roofline / rooflines   per-kernel CUDA-event durations (separate instrumented pass on the per-operator path: the same
             kernels the whole-step executor launches) against the measured peaks
  cpu_baseline           (N = 1) the reference's own CPU implementation on the host cores (bounded sample)

``--impl reference`` times the reference's CPU implementation (the live reference shipped under oracle/_ref when
present, else the oracle port; all host threads) on a bounded sample of the same workload and prints the same JSON
line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import hashlib
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
C3_STRUCT = 1024
C4_CELLS = 20
WORKLOAD_C2 = "C2: 256 x 108-atom FCC Cu (3x3x3 cells, a=3.615, jitter +-0.1 A), default M3GNet, energy+forces"
WORKLOAD_C3 = ("C3: 1024 MPF-like structures (20-200 atoms, 3-5 species, ragged triplet counts), default M3GNet, "
               "energy+forces, sharded by structure")
WORKLOAD_C4 = "C4: one 32 000-atom FCC Cu cell (20^3 cells, 72.3 A, jitter +-0.05 A), spatially domain-decomposed"
METRIC = "energy+forces atom-steps/sec"
UNIT = "atom-steps/s"
GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def dd_grid(world):
    """Near-cubic factorisation of the rank count (2 x 2 x 2 for 8, 3 x 2 x 1 for 6, ...)."""
    if world in GRIDS:
        return GRIDS[world]
    best = (world, 1, 1)
    for a in range(1, world + 1):
        if world % a:
            continue
        for b in range(1, world // a + 1):
            if (world // a) % b:
                continue
            g = tuple(sorted((a, b, world // a // b), reverse=True))
            if max(g) - min(g) < max(best) - min(best):
                best = g
    return best


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def source_hash():
    """Hash of the CUDA sources + ABI header: ties profiles/traffic.json to the build it was captured on."""
    h = hashlib.sha256()
    src = os.path.join(ROOT, "torch_m3gnet_b200", "csrc")
    for f in sorted(os.listdir(src)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(src, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "m3gnet_b200.h"), "rb").read())
    return h.hexdigest()[:16]


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
# workloads
# ----------------------------------------------------------------------------------------------------------
def pinned(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=dtype)).pin_memory()


class Workload:
    """Host-side description of a set of structures (+ pinned copies for the end-to-end leg)."""

    def __init__(self, lat, cart, z, sizes):
        self.lat = np.ascontiguousarray(lat, dtype=np.float64).reshape(-1, 3, 3)
        self.sizes = [int(s) for s in sizes]
        self.cart_pin = pinned(cart, np.float64)
        self.z_pin = pinned(z, np.int64)
        self.n_atoms = int(sum(self.sizes))

    def h2d_bytes(self):
        return int(self.cart_pin.numel() * 8 + self.z_pin.numel() * 8 + self.lat.size * 8 + 4 * (len(self.sizes) + 1))

    def build(self, device, want_triplet_index=True):
        import torch_m3gnet_b200 as m3g

        return m3g.Batch.from_arrays(self.lat, self.cart_pin, self.z_pin, self.sizes, HP["cutoff"],
                                     HP["threebody_cutoff"], device=device, want_triplet_index=want_triplet_index)


def c2_workload(seed0):
    from torch_m3gnet_b200 import synthetic

    return Workload(*synthetic.config2_batch(N_STRUCT, first_seed=seed0))


def _c3_header(s):
    rng = np.random.default_rng(1000 + s)
    n = int(rng.integers(20, 201))
    n_species = int(rng.integers(3, 6))
    rng.choice(np.arange(1, 95), size=n_species, replace=False)
    rho = rng.uniform(0.04, 0.09)
    a = (n / rho) ** (1.0 / 3.0)
    return n, a * (np.eye(3) + rng.uniform(-0.1, 0.1, size=(3, 3)) * (1 - np.eye(3)))


def _c3_make(s):
    from torch_m3gnet_b200 import synthetic

    return synthetic.mpf_like_structure(s)


def c3_workload(rank, world):
    """This rank's share of the 1 024 structures: every rank reads only the headers (atom count, cell) of all structures
    for the cost model, then generates the structures it owns (host worker pool; called BEFORE CUDA / NCCL are
    initialised, so that the pool can fork)."""
    from concurrent.futures import ProcessPoolExecutor

    from torch_m3gnet_b200 import shard

    heads = [_c3_header(s) for s in range(C3_STRUCT)]
    sizes = [h[0] for h in heads]
    assign, costs = shard.shard_structures(np.stack([h[1] for h in heads]), sizes, world, HP["cutoff"],
                                           HP["threebody_cutoff"])
    mine = assign[rank]
    t0 = time.time()
    workers = max(1, min(16, (os.cpu_count() or 2) // max(world, 1)))
    if workers > 1:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            structs = list(ex.map(_c3_make, mine, chunksize=8))
    else:
        structs = [_c3_make(s) for s in mine]
    gen_s = time.time() - t0
    arrays = (np.stack([st[0] for st in structs]), np.concatenate([st[1] for st in structs]),
              np.concatenate([st[2] for st in structs]), [len(st[1]) for st in structs])
    return arrays, dict(imbalance_predicted=shard.imbalance(costs, assign), generation_s=round(gen_s, 1),
                    structures=len(mine))


# ----------------------------------------------------------------------------------------------------------
# timing helpers
# ----------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, device):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.device = device
        self.on = self.world > 1
        if self.on:
            import torch.distributed as dist

            self.dist = dist
            dist.init_process_group("nccl", device_id=device)

    def sync(self):
        if self.on:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        if self.on:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        if self.on:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        if not self.on:
            return [float(v)]
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]


def time_resident(model, batch, steps, warmup, D, sampler=None):
    """Device-resident steps: CUDA events on the launching stream, barrier + synchronize on both sides."""
    from torch_m3gnet_b200 import _lib

    for _ in range(warmup):
        model(batch)
    D.sync()
    l0 = _lib.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = sampler if sampler is not None else _Null()
    with ctx:
        ev0.record()
        for _ in range(steps):
            model(batch)
        ev1.record()
        D.sync()
    ms = ev0.elapsed_time(ev1)
    return ms, (_lib.LAUNCHES - l0) // steps


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def time_e2e_positions(model, wl, device, steps, D):
    """coordinates / cells / atomic numbers in pinned host memory -> H2D -> GPU graph build -> model -> D2H of energies
    and forces; result of step k on the host before step k+1 starts.  Every copy is inside the timed region."""
    e_host = torch.empty(len(wl.sizes), dtype=torch.float32).pin_memory()
    f_host = torch.empty((wl.n_atoms, 3), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream()

    def run(n):
        for _ in range(n):
            b = wl.build(device, want_triplet_index=False)
            out = model(b)
            e_host.copy_(out["total_energy"], non_blocking=True)
            f_host.copy_(out["forces"], non_blocking=True)
            stream.synchronize()
            del b, out

    run(2)
    D.sync()
    t0 = time.perf_counter()
    run(steps)
    D.sync()
    dt = D.max(max(time.perf_counter() - t0, 1e-9))
    return dt, wl.h2d_bytes(), int(e_host.numel() * 4 + f_host.numel() * 4)


def time_e2e_graph_given(model, batch, device, steps, D):
    """A ready-made reference-format graph in pinned host memory (all nine tensors, int64 indices) -> H2D -> plan ->
    model -> D2H.  The upload + plan of step k+1 overlap the kernels of step k (side stream)."""
    import torch_m3gnet_b200 as m3g
    from torch_m3gnet_b200.data.material_graph import get_plan

    host = {k: batch[k].cpu().pin_memory() for k in ("pos", "atom_types", "num_triplet_i", "edge_index",
                                                     "edge_cell_shift", "num_triplet_ij", "triplet_edge_index",
                                                     "lattice", "batch")}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    n_atoms = host["pos"].shape[0]
    e_host = torch.empty(host["lattice"].shape[0], dtype=torch.float32).pin_memory()
    f_host = torch.empty((n_atoms, 3), dtype=torch.float32).pin_memory()
    main_stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=device)

    def stage_in():
        with torch.cuda.stream(side):
            dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
            b = m3g.Batch(pos=dev["pos"], atom_types=dev["atom_types"], num_triplet_i=dev["num_triplet_i"],
                          edge_index=dev["edge_index"], edge_cell_shift=dev["edge_cell_shift"],
                          num_triplet_ij=dev["num_triplet_ij"], triplet_edge_index=dev["triplet_edge_index"],
                          lattice=dev["lattice"])
            b["batch"] = dev["batch"]
            get_plan(b)
            ready = torch.cuda.Event()
            ready.record(side)
        return b, ready

    def run(n):
        nxt = stage_in() if n > 0 else None
        for k in range(n):
            b, ready = nxt
            main_stream.wait_event(ready)
            out = model(b)
            e_host.copy_(out["total_energy"], non_blocking=True)
            f_host.copy_(out["forces"], non_blocking=True)
            nxt = stage_in() if k + 1 < n else None
            main_stream.synchronize()
            del b, out

    run(2)
    D.sync()
    t0 = time.perf_counter()
    run(steps)
    D.sync()
    dt = D.max(max(time.perf_counter() - t0, 1e-9))
    return dt, int(h2d), int(e_host.numel() * 4 + f_host.numel() * 4)


# ----------------------------------------------------------------------------------------------------------
# per-kernel rooflines
# ----------------------------------------------------------------------------------------------------------
# ABI entry -> kernel name in the committed ncu launch list (profiles/traffic.json, written by tools/summarize_ncu.py)
TRAFFIC_KERNEL = {
    "conv_tc_bwd": "conv_tc_bwd2_kernel", "conv_tc_bwd_saved": "conv_tc_bwds_kernel",
    "conv_tc_fwd": "conv_tc4_fwd_kernel", "tb_atom_fwd": "tb_atom_fwd_kernel",
    "tb_atom_bwd": "tb_atom_bwd_kernel", "tb_mom_fwd": "tb_mom_fwd_kernel", "tb_mom_bwd": "tb_mom_bwd_kernel",
    "conv_gather_gz": "conv_gather_gz128_kernel",
    "segment_sum_add": "segment_sum_add_kernel", "tb_sigma_fwd": "tb_sigma_fwd_kernel",
    "tb_sigma_bwd": "tb_sigma_bwd_kernel<16>", "tb_radial": "tb_radial33_kernel",
    "tb_sigma64_fwd": "tb_sigma64_fwd_kernel", "tb_sigma64_bwd": "tb_sigma64_bwd_kernel",
    "segment_sum_parts": "segment_sum_parts_kernel", "tb_mom_red": "tb_mom_red_kernel",
    "tb_edge_update": "tb_edge_update_kernel<", "tb_mlp_adj": "tb_mlp_adj_kernel",
    "tb_mom_bwd_q": "tb_mom_bwd_kernel<",
}


def load_traffic():
    """Measured DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum of this bench command on B200,
    summarised by tools/summarize_ncu.py).  The file records the hash of the CUDA sources it was captured on; when the
    sources have changed since, the figures are NOT reported (traffic = null) rather than passed off as current."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return {}, "profiles/traffic.json absent"
    d = json.load(open(path))
    if d.get("source_hash") != source_hash():
        return {}, f"profiles/traffic.json was captured on another build ({d.get('source_hash')} vs {source_hash()})"
    return d.get("bytes_per_launch", {}), f"ncu capture of this build ({d.get('source_hash')}), {d.get('captured', '')}"


def kernel_model(E, T, N, Em, F=64, R=3):
    """Algorithmic bytes / flops per launch.  Em = member bonds (bonds inside the three-body cutoff).

    Gated-MLP pair: every byte the kernels are DESIGNED to move once — streams of e / e' / upstream rows, the 1 KB per
    bond of saved activations (written by the forward, read by the backward) and the 512 B per bond of first-layer
    adjoints that the per-atom gather consumes.  SURVEY.md §8(d)'s recompute-form figures (532 E + 512 N forward,
    800 E + 768 N backward, per conv = two MLPs) are reported next to them as `survey_bytes`."""
    mlp_mac = F * 2 * F + 2 * F * F
    single = {
        "conv_tc_fwd": ("hbm", 256 * E + 256 * E + 1024 * E + 12 * E + 8 * E + 512 * N, 2 * E * (mlp_mac + 2 * R * F),
                        (532 * E + 512 * N) // 2),
        "conv_tc_bwd_saved": ("hbm", 1024 * E + 256 * E + 256 * E + 512 * E + 12 * E + 4 * E + 12 * E + 128 * N,
                              2 * E * (4 * F * F), (800 * E + 768 * N) // 2),
        "conv_gather_gz": ("hbm", 512 * E + 4 * E + 4 * E + 1024 * N, 0, None),
        "segment_sum_add": ("hbm", 256 * E + 512 * N, 0, None),
    }
    groups = {
        # forward 4 T + 536 E + 36 N, backward 4 T + 332 E + 72 N (the kernels read no triplet index at
        # all; `moved` below is what they actually have to move: member bonds only for the per-bond three-body data)
        "threebody_fwd": (("tb_sigma64_fwd", "tb_sigma_fwd", "tb_mom_red", "tb_edge_update", "tb_mom_fwd", "tb_edge_basis_fwd", "tb_reduce_fwd", "tb_reduce_fwd_fast",
                           "tb_atom_fwd"), 4 * T + 536 * E + 36 * N, 512 * E + (16 + 36 + 36 + 4) * Em + 256 * N + 36 * N),
        "threebody_bwd": (("tb_mlp_adj", "tb_mom_bwd_q", "tb_mom_bwd", "tb_sigma64_bwd", "tb_sigma_bwd", "tb_gate_bwd", "tb_gate_bwd_fast", "tb_reduce_bwd",
                           "tb_reduce_bwd_sym", "tb_atom_bwd", "tb_edge_basis_bwd"), 4 * T + 332 * E + 72 * N,
                          256 * Em + (16 + 36 + 36 + 36 + 4) * Em + 16 * E + 36 * E + 36 * E + 256 * N + 72 * N),
    }
    return single, groups


def profile_pass(model, batch, steps, peaks):
    """Per-kernel durations with CUDA events around every ABI call (separate pass on the per-operator path: the
    whole-step executor launches the same kernels from C, where no per-launch events can be placed)."""
    from torch_m3gnet_b200 import _lib, engine

    old = engine.ENABLED
    engine.ENABLED = False
    try:
        model(batch)
        _lib.PROFILE = {}
        for _ in range(steps):
            model(batch)
        torch.cuda.synchronize()
        prof, _lib.PROFILE = _lib.PROFILE, None
    finally:
        engine.ENABLED = old
        _lib.PROFILE = None
    per = {}
    for name, evs in prof.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        per[name] = dict(calls_per_step=len(ms) / steps, ms_per_step=sum(ms) / steps, avg_ms=sum(ms) / len(ms))
    total = sum(v["ms_per_step"] for v in per.values())
    plan = batch._plan
    km, groups = kernel_model(plan.E, plan.T, plan.N, plan.n_members)
    tensor_peak = peaks["bf16_sustained"] or peaks["bf16"]
    traffic, traffic_note = load_traffic()
    rooflines = []
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        entry = dict(kernel="m3g_" + name, share=v["ms_per_step"] / total, avg_ms=v["avg_ms"],
                     launches_per_step=v["calls_per_step"], traffic=traffic.get(TRAFFIC_KERNEL.get(name, "")))
        if name in km:
            bound, nbytes, flops, survey = km[name]
            sec = v["avg_ms"] * 1e-3
            gbs = nbytes / sec / 1e9
            entry.update(bound=bound, achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                         algorithmic_bytes=nbytes)
            if flops:
                a = flops / sec / 1e12
                entry.update(tensor_tflops=a, tensor_frac=a / tensor_peak, tf32x3_frac=a / (tensor_peak / 2 / 3))
            if survey:
                entry.update(survey_bytes=survey, survey_frac=survey / sec / 1e9 / peaks["hbm"])
            if entry["traffic"]:
                entry["traffic_over_algorithmic"] = entry["traffic"] / nbytes
        rooflines.append(entry)
    for gname, (members, survey_bytes, moved) in groups.items():
        present = [m for m in members if m in per]
        if not present:
            continue
        sec = sum(per[m]["avg_ms"] for m in present) * 1e-3  # one op instance = one launch of each member kernel
        a = survey_bytes / sec / 1e9
        tr = [traffic.get(TRAFFIC_KERNEL.get(m, "")) for m in present]
        rooflines.append(dict(kernel=gname + " (" + "+".join("m3g_" + m for m in present) + ")",
                              traffic=(sum(tr) if all(t is not None for t in tr) else None),
                              share=sum(per[m]["ms_per_step"] for m in present) / total, avg_ms=sec * 1e3,
                              launches_per_step=per[present[0]]["calls_per_step"], bound="hbm", achieved=a,
                              peak=peaks["hbm"], unit="GB/s", frac=a / peaks["hbm"], algorithmic_bytes=survey_bytes,
                              moved_bytes=moved, moved_frac=moved / sec / 1e9 / peaks["hbm"],
                              triplets_per_s=plan.T / sec,
                              note="frac: SURVEY 8(d) formula (incl. 4 B per triplet that the index-free kernels do "
                                   "not read); moved_frac: the bytes the op has to move"))
    rooflines.sort(key=lambda r: -r["share"])
    return rooflines, total, traffic_note


# ----------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (live reference when shipped, else the oracle port)
# ----------------------------------------------------------------------------------------------------------
def reference_runner(sample):
    """Returns (callable running one energy+forces evaluation of the sample graphs, n_atoms, kind, description).
    sample: "c2" (2 of the 256 C2 structures) or "c3" (4 of the 1 024 C3 structures)."""
    from oracle import live_reference as lr
    from oracle import m3gnet_oracle as O
    from torch_m3gnet_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = O.HyperParams(**HP)
    if sample == "c2":
        structs = [synthetic.fcc_cu_supercell(3, 0.1, s) for s in range(2)]
        what = f"2 of the {N_STRUCT} C2 structures"
    else:
        structs = [synthetic.mpf_like_structure(s) for s in range(4)]
        what = f"4 of the {C3_STRUCT} C3 structures"
    g = O.collate([O.build_graph(lat, cart, z, HP["cutoff"], HP["threebody_cutoff"]) for lat, cart, z in structs])
    n_atoms = int(g["pos"].shape[0])
    torch.manual_seed(0)
    if lr.available():
        build_model, _, _ = lr.import_reference()
        model = build_model(HP["cutoff"], HP["threebody_cutoff"], HP["l_max"], HP["n_max"], HP["num_types"],
                            HP["embedding_dim"], HP["num_blocks"])

        def call():
            return model(lr.as_reference_graph(g))

        kind = "reference"
        desc = (f"{what} ({n_atoms} atoms) per call, graph given; the UNMODIFIED reference package "
                f"({lr.location()}) behind stand-ins for its absent wheels (torch CPU fp32, create_graph=True)")
    else:
        sd = O.init_params(hp, seed=0)
        fac = O.bessel_factors(hp.scaled_cutoff, hp.l_max, hp.n_max)

        def call():
            return O.forward(sd, hp, {k: v.clone() for k, v in g.items()}, factors=fac)

        kind = "port"
        desc = (f"{what} ({n_atoms} atoms) per call, graph given; oracle/m3gnet_oracle.py (torch CPU fp32, "
                f"create_graph=True) — the live reference is not shipped on this box")
    return call, n_atoms, kind, desc, cores


def cpu_baseline(reps=3):
    call, n_atoms, kind, desc, cores = reference_runner("c2")
    call()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        best = min(best, time.perf_counter() - t0)
    return dict(value=n_atoms / best, unit=UNIT, cores=cores, kind=kind, sample=desc + f"; best of {reps}",
                seconds_per_call=best)


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sample = "c2" if args.gpus <= 1 else "c3"
    call, n_atoms, kind, desc, cores = reference_runner(sample)
    for _ in range(args.warmup):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    value = n_atoms * args.steps / dt
    sample_txt = desc + "; atom-steps/s is per-atom so the sample scales linearly"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True,
                scaling="weak" if args.gpus <= 1 else "strong", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD_C2 if sample == "c2" else WORKLOAD_C3, sample=sample_txt,
                            note="one CPU process on rank 0's host cores regardless of --gpus"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample_txt),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# sharded configurations
# ----------------------------------------------------------------------------------------------------------
def run_c3(model, device, D, steps, warmup, prepared, sampler=None, with_e2e=True):
    arrays, info = prepared
    wl = Workload(*arrays)  # pinned host copies (needs the CUDA context of this rank's device)
    t0 = time.time()
    batch = wl.build(device)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    plan = batch._plan
    ms, launches = time_resident(model, batch, steps, warmup, D, sampler)
    per_rank = D.gather(ms / steps)
    ms_max = max(per_rank)
    n_tot = int(D.sum(plan.N))
    res = dict(workload=WORKLOAD_C3, value=n_tot / (ms_max * 1e-3), unit=UNIT, ms_per_step=ms_max, atoms=n_tot,
               bonds=int(D.sum(plan.E)), triplets=int(D.sum(plan.T)), structures=C3_STRUCT,
               per_rank_ms=[round(v, 3) for v in per_rank],
               imbalance_measured=ms_max / (sum(per_rank) / len(per_rank)),
               imbalance_predicted=info["imbalance_predicted"], generation_s=info["generation_s"],
               graph_build_s=round(build_s, 2), gpu_launches=int(launches),
               parallelism=f"structures LPT-sharded x{D.world}, no data-path collective")
    if with_e2e:
        dt, h2d, d2h = time_e2e_positions(model, wl, device, steps, D)
        res["e2e"] = dict(value=n_tot * steps / dt, unit=UNIT, h2d_bytes_per_step=int(D.sum(h2d)),
                          d2h_bytes_per_step=int(D.sum(d2h)))
    return res, batch


def run_dd(model, device, D, steps, warmup):
    """C4: one 32 000-atom cell.  N = 1: the undecomposed model (baseline of the curve).  N > 1: DomainStep on every
    rank (executor phases + NCCL halos), eager and as one CUDA graph; parity against the undecomposed model on rank 0."""
    import torch_m3gnet_b200 as m3g
    from torch_m3gnet_b200 import synthetic

    lat, cart, z = synthetic.fcc_cu_supercell(C4_CELLS, 0.05, 4)
    n = len(cart)
    res = dict(workload=WORKLOAD_C4, atoms=n, unit="atoms/s")
    # same architecture and cost as the headline model; weights x3 and an O(1) Bessel normalisation table make the
    # forces O(0.1 eV/A) and the three-body term visible, so that the parity figures below mean something
    sd3 = {k: (v.detach() * 3 if k.endswith("weight") else v.detach().clone()) for k, v in model.state_dict().items()}
    model = m3g.build_model(**HP, device=device)
    model.load_state_dict(sd3)
    fac = (torch.rand(3, 3, generator=torch.Generator().manual_seed(3)) + 0.5).to(device)
    for m in model.model:
        if hasattr(m, "nsb"):
            m.nsb.factors = fac
    full = None
    if D.rank == 0:
        full_batch = m3g.Batch.from_arrays(lat[None], cart, z, [n], 5.0, 4.0, device=device, want_triplet_index=False)
        for _ in range(max(warmup, 2)):
            full = model(full_batch)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            full = model(full_batch)
        ev1.record()
        torch.cuda.synchronize()
        res["single_gpu_ms_per_step"] = ev0.elapsed_time(ev1) / steps
        full = {"total_energy": full["total_energy"].clone(), "forces": full["forces"].clone()}
        del full_batch
    if not D.on:
        res.update(ms_per_step=res["single_gpu_ms_per_step"], value=n / (res["single_gpu_ms_per_step"] * 1e-3),
                   grid=[1, 1, 1], note="undecomposed model on one GPU (baseline of the domain-decomposition curve)")
        return res, False
    from torch_m3gnet_b200.domain import DomainBatch, DomainPlan, DomainStep

    single = D.max(res.get("single_gpu_ms_per_step", 0.0))
    plan = DomainPlan(lat, cart, z, dd_grid(D.world), 5.0)
    db = DomainBatch(plan, D.rank, 5.0, 4.0, device)

    def timed(step):
        for _ in range(max(warmup, 2)):
            out = step()
        D.sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            out = step()
        ev1.record()
        D.sync()
        return D.max(ev0.elapsed_time(ev1) / steps), out

    eager = DomainStep(model, db, capture=False)
    ms_eager, out = timed(eager)
    forces = torch.zeros((n, 3), device=device)
    forces[out["owned"]] = out["forces"]
    D.dist.all_reduce(forces)
    energy = out["total_energy"].clone()
    captured = False
    ms_graph = None
    variants = {"nccl_eager_phases": ms_eager}
    # halos as one pack-and-store kernel each into the peers' landing buffers (NVLink peer memory) + device barrier:
    # eager phases, then the whole step as one CUDA graph (kernels only: no collective call is left in the step)
    p2p_parity = None
    try:
        p2p = DomainStep(model, db, capture=False, exchange="p2p")
        variants["p2p_eager_phases"], out_p = timed(p2p)
        f_p = torch.zeros((n, 3), device=device)
        f_p[out_p["owned"]] = out_p["forces"]
        D.dist.all_reduce(f_p)
        p2p_parity = (float((out_p["total_energy"] - energy).abs().item() / n), float((f_p - forces).abs().max().item()))
        if os.environ.get("M3G_BENCH_DD_GRAPH", "1") != "0":
            p2p_graph = DomainStep(model, db, capture=True, exchange="p2p")
            variants["p2p_cuda_graph"], _ = timed(p2p_graph)
    except Exception as exc:  # peer access / symmetric memory unavailable: the NCCL path stands
        res["p2p_error"] = repr(exc)[:300]
    if os.environ.get("M3G_BENCH_DD_NCCL_GRAPH", "0") == "1":
        try:
            graphed = DomainStep(model, db, capture=True)
            variants["nccl_cuda_graph"], _ = timed(graphed)
            captured = True
        except Exception as exc:
            res["graph_capture_error"] = repr(exc)[:200]
    ms_graph = variants.get("p2p_cuda_graph", variants.get("nccl_cuda_graph"))
    ms_eager = min(v for k, v in variants.items() if k.endswith("eager_phases"))
    res["variants_ms_per_step"] = variants
    if p2p_parity is not None:
        res["p2p_vs_nccl_abs_dE_per_atom"], res["p2p_vs_nccl_max_abs_dF"] = p2p_parity
    best = min(ms_eager, ms_graph) if ms_graph is not None else ms_eager
    res.update(ms_per_step=best, value=n / (best * 1e-3), ms_per_step_eager_phases=ms_eager,
               ms_per_step_cuda_graph=ms_graph, single_gpu_ms_per_step=single,
               speedup_vs_single_gpu=(single / best if single else None),
               efficiency=(single / best / D.world if single else None), grid=list(dd_grid(D.world)),
               local_atoms_per_rank=[int(v) for v in D.gather(db.n_local)],
               owned_atoms_per_rank=[int(v) for v in D.gather(db.n_own)],
               ghosts_held=int(D.sum(db.n_local - db.n_own)), exchanges_per_step=eager.exchanges_per_step,
               halo_bytes_per_rank_per_exchange=int((db.n_local - db.n_own) * 64 * 4),
               parallelism="scheme B: one r_c ghost shell, bonds / triplets owned by their source atom, per-block halo "
                           "of ghost node features: NCCL all_to_all_single, or one pack-and-store kernel into the "
                           "peers' landing buffers over NVLink (symmetric memory) + device barrier")
    if D.rank == 0:
        res["abs_dE_per_atom"] = float((energy - full["total_energy"]).abs().item() / n)
        res["max_abs_dF"] = float((forces - full["forces"]).abs().max().item())
        res["max_abs_F"] = float(full["forces"].abs().max().item())
    return res, captured


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (used under ncu only)")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the single-GPU C3 / C4 figures")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_env, rank_env = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    c3_prepared = None
    if world_env > 1 or not args.no_extra:
        c3_prepared = c3_workload(rank_env, world_env)  # host-only; forks a worker pool, hence before any CUDA call
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    D = Dist(device)

    import torch_m3gnet_b200 as m3g

    peaks = load_peaks()
    torch.manual_seed(0)
    model = m3g.build_model(**HP, device=None)  # CPU init (seed 0) so that every rank holds the same weights
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = m3g.build_model(**HP, device=device)
    model.load_state_dict(sd_cpu)
    captured = False

    if not D.on:
        # ---------------- N = 1: C2 headline ----------------
        wl = c2_workload(0)
        t0 = time.time()
        batch = wl.build(device)
        torch.cuda.synchronize()
        build_s = time.time() - t0
        plan = batch._plan
        with ClockSampler(local_rank) as clocks:
            ms, launches = time_resident(model, batch, args.steps, args.warmup, D)
        value = plan.N * args.steps / (ms * 1e-3)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload=WORKLOAD_C2, atoms_per_gpu=plan.N, edges_per_gpu=plan.E,
                                triplets_per_gpu=plan.T, structures_per_gpu=N_STRUCT, parallelism="single GPU",
                                l2="per-step working set (edge features 6 x 297 MB + saved activations) >> 126 MB L2",
                                graph_build_s=build_s),
                    triplets_per_s=plan.T * args.steps / (ms * 1e-3), clocks=clocks.summary(),
                    gpu_launches=int(launches))
        if not args.no_e2e:
            dt, h2d, d2h = time_e2e_positions(model, wl, device, args.steps, D)
            line["e2e"] = dict(value=plan.N * args.steps / dt, unit=UNIT, h2d_bytes_per_step=h2d,
                               d2h_bytes_per_step=d2h,
                               note="pinned host coordinates / cells / atomic numbers -> H2D -> neighbour list + "
                                    "triplets on the GPU -> model -> D2H of energies + forces, every step")
            dt, h2d, d2h = time_e2e_graph_given(model, batch, device, args.steps, D)
            line["e2e_graph_given"] = dict(value=plan.N * args.steps / dt, unit=UNIT, h2d_bytes_per_step=h2d,
                                           d2h_bytes_per_step=d2h,
                                           note="ready-made reference-format graph (nine tensors, int64 indices incl. "
                                                "the (2,T) triplet list) uploaded from pinned host memory every step")
        if not args.no_profile:
            rooflines, kernel_ms, traffic_note = profile_pass(model, batch, 2, peaks)
            dominant = next((r for r in rooflines if "bound" in r), None)
            if dominant is not None:
                line["roofline"] = dict(bound=dominant["bound"], achieved=dominant["achieved"], peak=dominant["peak"],
                                        unit=dominant["unit"], frac=dominant["frac"], traffic=dominant.get("traffic"),
                                        kernel=dominant["kernel"],
                                        peak_source=peaks["source"] + " (MEASURED_PEAKS.json)",
                                        note="dominant kernel of the step; algorithmic bytes (DESIGN.md 4) / CUDA-event "
                                             "duration against the measured HBM copy peak; traffic: " + traffic_note)
            line["rooflines"] = rooflines[:26]
            line["kernel_ms_per_step_operator_path"] = kernel_ms
        del batch
        if not args.no_extra:
            c3, b3 = run_c3(model, device, D, max(args.steps // 2, 3), 3, c3_prepared, with_e2e=not args.no_e2e)
            line["c3"] = c3
            del b3
            dd, _ = run_dd(model, device, D, max(args.steps // 2, 3), 3)
            line["dd"] = dd
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
        return

    # ---------------- N > 1: C3 strong scaling headline + C4 domain decomposition ----------------
    with ClockSampler(local_rank) as clocks:
        c3, batch = run_c3(model, device, D, args.steps, args.warmup, c3_prepared, with_e2e=not args.no_e2e)
    line = dict(metric=METRIC, value=c3["value"], unit=UNIT, n_gpus=D.world, steps=args.steps, warmup=args.warmup,
                ms_per_step=c3["ms_per_step"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload=WORKLOAD_C3, atoms=c3["atoms"], bonds=c3["bonds"], triplets=c3["triplets"],
                            structures=C3_STRUCT, parallelism=c3["parallelism"], per_rank_ms=c3["per_rank_ms"],
                            imbalance_measured=c3["imbalance_measured"], imbalance_predicted=c3["imbalance_predicted"],
                            l2="per-rank working set >> 126 MB L2 (edge features and saved activations)",
                            note="total work fixed (strong scaling); the N = 1 line reports the same workload on one "
                                 "GPU under \"c3\""),
                clocks=clocks.summary(), gpu_launches=c3["gpu_launches"], c3=c3)
    if "e2e" in c3:
        line["e2e"] = c3["e2e"]
    if not args.no_profile and D.rank == 0:
        rooflines, kernel_ms, traffic_note = profile_pass(model, batch, 2, peaks)
        dominant = next((r for r in rooflines if "bound" in r), None)
        if dominant is not None:
            line["roofline"] = dict(bound=dominant["bound"], achieved=dominant["achieved"], peak=dominant["peak"],
                                    unit=dominant["unit"], frac=dominant["frac"], traffic=None,
                                    kernel=dominant["kernel"], peak_source=peaks["source"] + " (MEASURED_PEAKS.json)",
                                    note="rank 0's shard; ncu traffic is captured at N = 1 only")
        line["rooflines"] = rooflines[:8]
    del batch
    D.sync()
    dd, captured = run_dd(model, device, D, args.steps, args.warmup)
    line["dd"] = dd
    D.sync()
    if D.rank == 0:
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    if captured:
        # collectives captured in a CUDA graph leave work objects the NCCL watchdog never sees complete; tearing the
        # process group down would wait for them.  Everything is reported: leave through process exit.
        D.sync()
        os._exit(0)
    D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
