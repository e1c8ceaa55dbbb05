"""CPU, world_size 2 over gloo: host-side logic of the two multi-GPU modes (structure sharding, halo exchange)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import m3gnet_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(fn, world, port) + args, nprocs=world, join=True)


def _entry(rank, fn, world, port, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
def test_structure_assignment_is_a_balanced_partition():
    from torch_m3gnet_b200 import shard

    structs = [O.mpf_like_structure(s) for s in range(64)]
    lats = np.stack([s[0] for s in structs])
    sizes = [len(s[1]) for s in structs]
    for world in (1, 2, 4, 8):
        assign, costs = shard.shard_structures(lats, sizes, world, 5.0, 4.0)
        flat = sorted(i for a in assign for i in a)
        assert flat == list(range(64))
        assert shard.imbalance(costs, assign) < 1.10, (world, shard.imbalance(costs, assign))
    # the cost model tracks the real edge / triplet counts of the oracle's graphs
    real = []
    for lat, cart, z in structs[:12]:
        g = O.build_graph(lat, cart, z, 5.0, 4.0)
        real.append(g["edge_index"].shape[1] + 0.05 * g["triplet_edge_index"].shape[1])
    pred = costs[:12]
    assert np.corrcoef(real, pred)[0, 1] > 0.97


def _gather_worker(rank, world):
    from torch_m3gnet_b200 import shard

    costs = [float(c) for c in np.random.default_rng(3).uniform(1, 10, size=11)]
    assign = shard.assign_structures(costs, world)
    local = torch.tensor([[float(i), 2.0 * i] for i in assign[rank]]).reshape(-1, 2)
    out = shard.gather_by_structure(local, assign, rank, world)
    want = torch.tensor([[float(i), 2.0 * i] for i in range(11)])
    assert torch.equal(out, want)


def test_gather_by_structure_gloo():
    _run(_gather_worker, 2)


def _halo_worker(rank, world, grid):
    from torch_m3gnet_b200.domain import DistHaloFn, DomainBatch, DomainPlan

    lat, cart, z = O.fcc_supercell(5, jitter=0.1, seed=2)
    cart = cart + np.array([0.3, -7.0, 25.0])  # unwrapped input coordinates
    plan = DomainPlan(lat, cart, z, grid, 5.0)
    assert plan.world == world
    d = DomainBatch.exchange_only(plan, rank)
    W = 5
    table = torch.from_numpy(np.random.default_rng(0).normal(size=(len(cart), W))).float()
    x = torch.zeros(d.n_local, W)
    x[: d.n_own] = table[d.global_owned]
    x.requires_grad_(True)
    out = DistHaloFn.apply(x, d, None)
    ghost_atoms = torch.as_tensor(plan.ghost_atom[rank], dtype=torch.long)
    assert torch.equal(out[: d.n_own], table[d.global_owned])
    assert torch.equal(out[d.n_own:], table[ghost_atoms])
    # backward: every ghost copy's gradient comes home to its owner
    coeff = torch.from_numpy(np.random.default_rng(10 + rank).normal(size=(d.n_local, W))).float()
    (out * coeff).sum().backward()
    contrib = torch.zeros(len(cart), W)
    contrib.index_add_(0, d.global_owned, coeff[: d.n_own])
    contrib.index_add_(0, ghost_atoms, coeff[d.n_own:])
    dist.all_reduce(contrib)
    mine = torch.zeros(len(cart), W)
    mine.index_add_(0, d.global_owned, x.grad[: d.n_own])
    dist.all_reduce(mine)
    torch.testing.assert_close(mine, contrib)
    assert torch.all(x.grad[d.n_own:] == 0)


def test_halo_exchange_gloo_slab():
    _run(_halo_worker, 2, (2, 1, 1))


def test_halo_exchange_gloo_other_axis():
    _run(_halo_worker, 2, (1, 1, 2))


def test_domain_plan_geometry():
    """Every neighbour (within r_c) of an owned atom is present locally, as an owned atom or as a ghost image."""
    from torch_m3gnet_b200.domain import DomainPlan

    lat, cart, z = O.fcc_supercell(4, jitter=0.1, seed=5)
    shear = np.eye(3) + 0.08 * np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0.0]])
    lat, cart = lat @ shear, cart @ shear
    src, dst, img, dist_ = O.neighbor_list_bruteforce(lat, cart, 5.0)
    for grid in ((2, 1, 1), (2, 2, 1), (2, 2, 2)):
        plan = DomainPlan(lat, cart, z, grid, 5.0)
        assert sorted(np.concatenate(plan.owned).tolist()) == list(range(len(cart)))
        for r in range(plan.world):
            pos, zz, n_own = plan.local_arrays(r)
            own = plan.owned[r]
            # local neighbour counts of owned atoms (non-periodic cloud) must equal the periodic counts
            d2 = ((pos[:n_own, None, :] - pos[None, :, :]) ** 2).sum(-1)
            local_counts = ((d2 < 25.0 + 1e-8) & (d2 > 1e-12)).sum(1)
            want = np.bincount(src, minlength=len(cart))[own]
            assert np.array_equal(local_counts, want), (grid, r)
    with pytest.raises(ValueError):
        DomainPlan(lat, cart, z, (4, 1, 1), 5.0)  # 3.6 Å slabs are thinner than the cutoff
