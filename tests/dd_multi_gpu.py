"""torchrun entry: one process per GPU, NCCL halo exchange; rank 0 compares against the undecomposed model.

  python -m torch.distributed.run --nproc-per-node N tests/dd_multi_gpu.py [--reps 8] [--cells 8]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torch_m3gnet_b200 import build_model, synthetic  # noqa: E402
from torch_m3gnet_b200.data.material_graph import Batch  # noqa: E402
from torch_m3gnet_b200.domain import DomainBatch, DomainPlan, DomainStep, evaluate_distributed  # noqa: E402

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=6)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mode", default="engine", choices=["autograd", "engine", "graph", "p2p", "p2p-graph"],
                    help="autograd: per-operator Functions + DistHaloFn; engine: m3g_step_run phases + exchanges; "
                         "graph: the engine step captured in a CUDA graph (NCCL inside)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)
    cpu_model = build_model(5.0, 4.0, 3, 3, 95, 64, 3)
    sd = {k: (v * 3 if k.endswith("weight") else v) for k, v in cpu_model.state_dict().items()}
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    model.load_state_dict(sd)
    fac = (torch.rand(3, 3, generator=torch.Generator().manual_seed(3)) + 0.5).to(device)
    for i in (6, 8, 10):
        model.model[i].nsb.factors = fac
    lat, cart, z = synthetic.fcc_cu_supercell(args.cells, 0.05, 4)
    plan = DomainPlan(lat, cart, z, GRIDS[world], 5.0)
    db = DomainBatch(plan, rank, 5.0, 4.0, device)
    if args.mode == "autograd":
        step = lambda: evaluate_distributed(model, db)  # noqa: E731
    else:
        step = DomainStep(model, db, capture=args.mode.endswith("graph"),
                          exchange="p2p" if args.mode.startswith("p2p") else "nccl")
    res = step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step()
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / args.steps
    res = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in res.items()}
    # gather forces to rank 0
    n = len(cart)
    forces = torch.zeros((n, 3), device=device)
    forces[res["owned"]] = res["forces"]
    dist.all_reduce(forces)
    if rank == 0:
        full = model(Batch.from_arrays(lat[None], cart, z, [n], 5.0, 4.0, device=device))
        dE = (res["total_energy"] - full["total_energy"]).abs().item() / n
        dF = (forces - full["forces"]).abs().max().item()
        fmax = full["forces"].abs().max().item()
        print(f"[dd-nccl] mode={args.mode} world={world} atoms={n} local={db.n_local} (owned {db.n_own}) step={dt * 1e3:.2f} ms "
              f"atoms/s={n / dt:.3e} |dE|/atom={dE:.3e} max|dF|={dF:.3e} max|F|={fmax:.3e}")
        assert dE <= 1e-6 and dF <= 1e-5 + 1e-4 * fmax
        print("DD-OK", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    if args.mode == "graph":
        # collectives captured in a CUDA graph leave work objects the NCCL watchdog never sees complete:
        # destroy_process_group would wait for them; results are out, leave through process exit
        sys.stdout.flush()
        os._exit(0)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
