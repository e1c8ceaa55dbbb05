"""BASELINE config 3: MPF-like ragged multi-element batch (1024 structures, 20-200 atoms, 3-5 species), sharded BY
STRUCTURE with the longest-processing-time cost model of torch_m3gnet_b200/shard.py; no data-path collective.

  python tools/bench_c3.py [--structures 1024] [--steps 5]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c3.py

Every rank reads only the header (atom count, cell) of all structures for the cost model and builds the structures it
owns.  Prints one JSON line on rank 0: whole-job atom-steps/s (max time over ranks), predicted and measured imbalance.
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
from torch_m3gnet_b200 import shard, synthetic  # noqa: E402


def header(s: int):
    """(n_atoms, lattice) of structure s: the first random draws of synthetic.mpf_like_structure, nothing else."""
    rng = np.random.default_rng(1000 + s)
    n = int(rng.integers(20, 201))
    n_species = int(rng.integers(3, 6))
    rng.choice(np.arange(1, 95), size=n_species, replace=False)
    rho = rng.uniform(0.04, 0.09)
    a = (n / rho) ** (1.0 / 3.0)
    return n, a * (np.eye(3) + rng.uniform(-0.1, 0.1, size=(3, 3)) * (1 - np.eye(3)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--structures", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    heads = [header(s) for s in range(args.structures)]
    sizes = [h[0] for h in heads]
    assign, costs = shard.shard_structures(np.stack([h[1] for h in heads]), sizes, world, 5.0, 4.0)
    mine = assign[rank]
    t0 = time.time()
    structs = [synthetic.mpf_like_structure(s) for s in mine]
    gen_s = time.time() - t0
    assert all(len(st[1]) == sizes[s] for st, s in zip(structs, mine))
    batch = m3g.Batch.from_arrays(np.stack([st[0] for st in structs]), np.concatenate([st[1] for st in structs]),
                                  np.concatenate([st[2] for st in structs]), [len(st[1]) for st in structs], 5.0, 4.0,
                                  device=dev)
    torch.manual_seed(0)
    cpu_model = m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3)
    model = m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=dev)
    model.load_state_dict(cpu_model.state_dict())
    plan = batch._plan
    for _ in range(args.warmup):
        model(batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = model(batch)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    stats = torch.tensor([ms, plan.N, plan.E, plan.T], dtype=torch.float64, device=dev)
    if world > 1:
        allst = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
    else:
        allst = [stats]
    if rank == 0:
        tab = torch.stack(allst).cpu().numpy()
        ms_max, ms_mean = float(tab[:, 0].max()), float(tab[:, 0].mean())
        n_tot = int(tab[:, 1].sum())
        print(json.dumps(dict(
            metric="energy+forces atom-steps/sec", value=n_tot / (ms_max * 1e-3), unit="atom-steps/s", n_gpus=world,
            ms_per_step=ms_max, steps=args.steps, dtype="f32", data="synthetic", scaling="strong",
            config=dict(workload=f"C3: {args.structures} MPF-like structures (20-200 atoms, 3-5 species), sharded by "
                                 "structure (LPT on predicted cost), no data-path collective",
                        atoms=n_tot, bonds=int(tab[:, 2].sum()), triplets=int(tab[:, 3].sum()),
                        per_rank_ms=[round(float(v), 3) for v in tab[:, 0]],
                        per_rank_atoms=[int(v) for v in tab[:, 1]], generation_s=round(gen_s, 1)),
            imbalance_predicted=shard.imbalance(costs, assign), imbalance_measured=ms_max / ms_mean,
            energy_check=float(out["total_energy"].abs().sum().item()))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
