"""GPU: spatial domain decomposition — all ranks emulated in lockstep on one GPU (same plan and exchange order as
the NCCL path) must reproduce the undecomposed periodic model; plus the real multi-process path when >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import m3gnet_oracle as O
from tests.util import golden, report, state_dict_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(device):
    from torch_m3gnet_b200 import build_model

    sd = {k: (v * 3 if k.endswith("weight") else v) for k, v in state_dict_of(golden("c1_default")).items()}
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    model.load_state_dict(sd)
    # an O(1) Bessel table so that the three-body term (and its ghost-x dependence) really matters
    torch.manual_seed(3)
    fac = (torch.rand(3, 3) + 0.5).to(device)
    for i in (6, 8, 10):
        model.model[i].nsb.factors = fac
    return model


@pytest.mark.parametrize("offset", [False, True])
@pytest.mark.parametrize("grid", [(2, 1, 1), (1, 2, 2), (2, 2, 2)])
def test_emulated_domains_match_full_cell(device, grid, offset):
    from torch_m3gnet_b200.data.material_graph import Batch
    from torch_m3gnet_b200.domain import DomainBatch, DomainPlan, evaluate_emulated

    lat, cart, z = O.fcc_supercell(6, jitter=0.05, seed=4)  # 864 atoms, 21.7 Å box
    z = z.copy()
    z[::3] = 13  # two species
    if offset:
        # unwrapped input coordinates.  The undecomposed model then works with float32 positions up to 77 Å
        # (ulp 7.6e-6 Å) while the domains use wrapped (compact) positions: the comparison tolerance reflects the
        # float32 rounding of the *inputs*, not of the decomposition (the offset=False case is tight).
        cart = cart + np.array([1.0, -30.0, 55.0])
    model = _model(device)
    full = model(Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 4.0, device=device))
    plan = DomainPlan(lat, cart, z, grid, 5.0)
    dbs = [DomainBatch(plan, r, 5.0, 4.0, device) for r in range(plan.world)]
    res = evaluate_emulated(model, dbs)
    n = len(cart)
    ghosts = sum(len(g) for g in plan.ghost_atom)
    print(f"[dd] grid={grid} owned={n} ghosts={ghosts} (x{1 + ghosts / n:.2f} atoms held)")
    dE = (res["total_energy"] - full["total_energy"]).abs().item() / n
    print(f"[dd] |dE|/atom = {dE:.3e}")
    assert dE <= 1e-6
    report(f"dd {grid} offset={offset} forces", res["forces"], full["forces"], 4e-5 if offset else 2e-6, 1e-5)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_distributed_domains_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dd_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DD-OK" in r.stdout


def test_domain_step_single_rank_process_group(device):
    """DomainStep (whole-step executor phases + NCCL halo exchanges) with a one-rank process group: every ghost is a
    periodic image of an owned atom, so all halos are self-exchanges; against the undecomposed model."""
    import torch.distributed as dist

    from torch_m3gnet_b200.data.material_graph import Batch
    from torch_m3gnet_b200.domain import DomainBatch, DomainPlan, DomainStep

    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29531", world_size=1, rank=0,
                                device_id=device)
        created = True
    try:
        lat, cart, z = O.fcc_supercell(5, jitter=0.05, seed=4)  # 500 atoms, 18.1 A box
        z = z.copy()
        z[::3] = 13
        model = _model(device)
        full = model(Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 4.0, device=device))
        plan = DomainPlan(lat, cart, z, (1, 1, 1), 5.0)
        db = DomainBatch(plan, 0, 5.0, 4.0, device)
        n = len(cart)
        step = DomainStep(model, db, capture=False)
        for _ in range(2):
            res = step()
        forces = torch.zeros((n, 3), device=device)
        forces[res["owned"]] = res["forces"]
        dE = (res["total_energy"] - full["total_energy"]).abs().item() / n
        print(f"[dd-step] local={db.n_local} owned={db.n_own} |dE|/atom={dE:.3e} exchanges/step={step.exchanges_per_step}")
        assert dE <= 1e-6
        report("dd-step forces", forces, full["forces"], 2e-6, 1e-5)
    finally:
        if created:
            torch.cuda.synchronize()
            dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["engine", "graph", "p2p", "p2p-graph"])
def test_domain_step_one_rank_subprocess(mode):
    """The torchrun entry with one rank: executor phases with NCCL halos, the whole step (NCCL exchanges included)
    captured in a CUDA graph, and the same two with the halos as pack-and-store kernels into symmetric-memory landing
    buffers + device barriers.  Runs in its own process: a process group that has captured collectives is left to
    process exit."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "dd_multi_gpu.py"), "--cells", "5",
           "--mode", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DD-OK" in r.stdout
