"""GPU: Verlet (skin) neighbour list and the trajectory-side callers.

The bar is the graph builder's own: every frame's Batch must be bit-identical (index tensors, image shifts, triplets,
build distances) to ``Batch.from_arrays`` on the same coordinates -- which tests/test_gpu_graph.py pins to the oracle
and the live-reference fixtures -- and the model outputs on it identical to those on the fresh graph."""
import numpy as np
import pytest
import torch

from torch_m3gnet_b200 import Batch, Fire, M3GNetCalculator, VelocityVerlet, VerletList, build_model, synthetic

pytestmark = pytest.mark.gpu

INDEX_KEYS = ("edge_index", "edge_cell_shift", "num_triplet_i", "num_triplet_ij", "triplet_edge_index", "atom_types",
              "batch", "pos", "lattice")


def _same_graph(a, b):
    for k in INDEX_KEYS:
        assert torch.equal(a[k], b[k]), k
    assert a["num_edges"] == b["num_edges"] and a["num_triplets"] == b["num_triplets"]
    assert torch.equal(a._private["edge_distances_build"], b._private["edge_distances_build"])
    pa, pb = a._plan, b._plan
    for k in ("edge_ptr", "tri_ptr", "tri_e2", "in_ptr", "member_edges"):
        assert torch.equal(getattr(pa, k), getattr(pb, k)), k
    assert torch.equal(pa.in_perm[:pa.E], pb.in_perm[:pb.E])  # allocated with one spare slot when there are no bonds
    assert (pa.N, pa.E, pa.T, pa.B, pa.tri_dense, pa.max_members) == (pb.N, pb.E, pb.T, pb.B, pb.tri_dense,
                                                                      pb.max_members)


def _cases():
    lat1, cart1, z1 = synthetic.fcc_cu_supercell(2, 0.05, 1)                    # C1: 32 atoms, images on every bond
    lat6, cart6, z6 = synthetic.fcc_cu_supercell(6, 0.1, 2)                     # 864 atoms, 21.7 A: cell-list path
    tiny = (np.array([[2.6, 0, 0], [0.3, 2.7, 0], [0.1, -0.2, 2.9]]), np.array([[0.1, 0.2, 0.3]]), np.array([29]))
    rag = [synthetic.mpf_like_structure(s) for s in range(5)]                   # ragged multi-species batch
    far = (np.eye(3) * 20.0, np.array([[1.0, 2.0, 3.0], [11.0, 12.0, 13.0]]), np.array([29, 8]))  # no bonds at all
    return {
        "isolated": ([far[0]], far[1], far[2], [2]),
        "c1": ([lat1], cart1, z1, [len(cart1)]),
        "cells864": ([lat6], cart6, z6, [len(cart6)]),
        "self_images": ([tiny[0]], tiny[1], tiny[2], [1]),
        "ragged5": ([r[0] for r in rag], np.concatenate([r[1] for r in rag]), np.concatenate([r[2] for r in rag]),
                    [len(r[1]) for r in rag]),
    }


@pytest.mark.parametrize("name", ["c1", "cells864", "self_images", "ragged5", "isolated"])
def test_verlet_frames_equal_fresh_builds(device, name):
    lats, cart, z, sizes = _cases()[name]
    lats = np.stack(lats)
    vl = VerletList(lats, z, sizes, 5.0, 4.0, skin=0.5, device=device, want_triplet_index=True)
    rng = np.random.default_rng(7)
    cart = cart.copy()
    rebuilds_seen = []
    for frame in range(14):
        got = vl.update(cart)
        want = Batch.from_arrays(lats, cart, z, sizes, 5.0, 4.0, device=device)
        _same_graph(got, want)
        rebuilds_seen.append(vl.n_rebuilds)
        cart = cart + rng.normal(0.0, 0.04, size=cart.shape)   # random walk: ~0.25 A after 13 frames -> crosses skin/2
    assert vl.n_frames == 14
    assert 2 <= vl.n_rebuilds < 14, rebuilds_seen                # reused between rebuilds, rebuilt when needed
    # a device-resident float64 tensor is accepted as well and gives the same frame
    got = vl.update(torch.as_tensor(cart).to(device))
    _same_graph(got, Batch.from_arrays(lats, cart, z, sizes, 5.0, 4.0, device=device))


def test_verlet_rebuild_trigger_and_lattice_change(device):
    lat, cart, z = synthetic.fcc_cu_supercell(2, 0.05, 3)
    vl = VerletList(lat[None], z, [len(cart)], 5.0, 4.0, skin=0.4, device=device)
    vl.update(cart)
    assert vl.n_rebuilds == 1
    c2 = cart.copy()
    c2[5] += np.array([0.19, 0.0, 0.0])        # below skin/2 = 0.2
    vl.update(c2)
    assert vl.n_rebuilds == 1
    c2[5] += np.array([0.02, 0.0, 0.0])        # 0.21 from the reference frame
    vl.update(c2)
    assert vl.n_rebuilds == 2
    vl.update(c2 + 5.0)                        # rigid translation by more than the skin: rebuild, same bonds
    assert vl.n_rebuilds == 3
    lat2 = lat * 1.01
    vl.set_lattice(lat2[None])
    got = vl.update(c2)
    assert vl.n_rebuilds == 4
    want = Batch.from_arrays(lat2[None], c2, z, [len(c2)], 5.0, 4.0, device=device, want_triplet_index=False)
    for k in ("edge_index", "edge_cell_shift", "num_triplet_ij"):
        assert torch.equal(got[k], want[k])
    bad = c2.copy()
    bad[0, 0] = np.nan                         # NaN coordinates must not silently reuse the candidates
    vl.update(bad)
    assert vl.n_rebuilds == 5


def _model(device):
    torch.manual_seed(11)
    model = build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=device)
    sd = model.state_dict()
    model.load_state_dict({k: (v * 3 if k.endswith("weight") else v) for k, v in sd.items()})  # O(0.1) eV/A forces
    return model


class _Atoms:
    """The part of ase.Atoms the calculator protocol touches."""

    def __init__(self, cell, positions, numbers):
        self.cell, self.positions, self.numbers = cell, positions, numbers

    def get_cell(self):
        return self.cell

    def get_positions(self):
        return self.positions

    def get_atomic_numbers(self):
        return self.numbers


def test_calculator_equals_fresh_graph_per_frame(device):
    lat, cart, z = synthetic.fcc_cu_supercell(2, 0.05, 4)
    model = _model(device)
    calc = M3GNetCalculator(model, 5.0, 4.0, skin=0.5, device=device)
    rng = np.random.default_rng(0)
    for frame in range(6):
        e, f, s = calc.compute(lat, cart, z)
        out = model(Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 4.0, device=device))
        assert torch.equal(e, out["total_energy"]) and torch.equal(f, out["forces"]) and torch.equal(s, out["stresses"])
        cart = cart + rng.normal(0.0, 0.03, size=cart.shape)
    assert calc.neighbor_list.n_rebuilds < calc.neighbor_list.n_frames
    atoms = _Atoms(lat, cart, z)
    res = calc.calculate(atoms, ("energy", "forces"))
    out = model(Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 4.0, device=device))
    assert res["energy"] == float(out["total_energy"][0]) == res["free_energy"]
    np.testing.assert_array_equal(res["forces"], out["forces"].double().cpu().numpy())
    np.testing.assert_array_equal(res["m3gnet_stresses"], out["stresses"][0].double().cpu().numpy())
    assert calc.get_forces(atoms).shape == (32, 3) and calc.get_reference_stresses(atoms).shape == (6,)
    assert calc.calculate(atoms, ("forces",)) is res  # unchanged atoms: cached results, no second evaluation
    with pytest.raises(NotImplementedError):
        calc.calculate(atoms, ("stress",))  # the reference's row is not ASE's stress: not advertised
    assert isinstance(calc.get_potential_energy(atoms), float)
    with pytest.raises(NotImplementedError):
        calc.calculate(atoms, ("magmoms",))
    # another composition: the candidate list is replaced, not reused
    lat2, cart2, z2 = synthetic.mpf_like_structure(3)
    e2, f2, _ = calc.compute(lat2, cart2, z2)
    out2 = model(Batch.from_arrays(lat2[None], cart2, z2, [len(cart2)], 5.0, 4.0, device=device))
    assert torch.equal(e2, out2["total_energy"]) and torch.equal(f2, out2["forces"])


def test_velocity_verlet_driver(device):
    lat, cart, z = synthetic.fcc_cu_supercell(2, 0.05, 5)
    model = _model(device)
    calc = M3GNetCalculator(model, 5.0, 4.0, skin=0.5, device=device)
    rng = np.random.default_rng(1)
    v0 = rng.normal(0.0, 0.003, size=cart.shape)   # ~ 300 K for Cu in A/fs
    v0 -= v0.mean(axis=0)
    md = VelocityVerlet(calc, lat, cart, z, np.full(len(cart), 63.546), dt=1.0, velocities=v0)
    p0 = md.pos.clone()
    ke0 = md.kinetic_energy()
    assert 0.5 < ke0 / (1.5 * 32 * 8.617e-5 * 300.0) < 2.0        # sanity of the unit conversion
    md.step(20)
    assert md.n_steps == 20 and torch.isfinite(md.forces).all() and np.isfinite(md.potential_energy())
    assert (md.pos - p0).abs().max().item() > 0.02
    # translation invariance of the energy <=> forces sum to zero <=> total momentum is conserved
    mom = (md.mass.unsqueeze(1) * md.vel).sum(0).abs().max().item()
    assert mom < 1e-3 * (md.mass.unsqueeze(1) * md.vel).abs().sum().item()
    # the driver's last frame equals a fresh evaluation at its coordinates
    out = model(Batch.from_arrays(lat[None], md.pos.cpu().numpy(), z, [32], 5.0, 4.0, device=device))
    assert torch.equal(out["forces"], md.forces)
    assert calc.neighbor_list.n_frames == 21 and calc.neighbor_list.n_rebuilds <= 3


def test_fire_relaxation_lowers_energy_and_forces(device):
    lat, cart, z = synthetic.fcc_cu_supercell(2, 0.08, 6)
    model = _model(device)
    calc = M3GNetCalculator(model, 5.0, 4.0, skin=0.5, device=device)
    opt = Fire(calc, lat, cart, z)
    e0, f0 = float(opt.energy[0]), opt.fmax()
    opt.run(fmax=0.0, steps=60)
    e1, f1 = float(opt.energy[0]), opt.fmax()
    print(f"[fire] E {e0:.6f} -> {e1:.6f} eV, fmax {f0:.4f} -> {f1:.4f} eV/A, {calc.neighbor_list.n_rebuilds} rebuilds")
    assert opt.n_steps == 60 and e1 < e0 and f1 < f0
    out = model(Batch.from_arrays(lat[None], opt.pos.cpu().numpy(), z, [32], 5.0, 4.0, device=device))
    assert torch.equal(out["forces"], opt.forces)


def test_graph_replay_while_bonds_are_unchanged(device):
    """graph_replay=True: frames with an unchanged bond set are replayed from a CUDA graph and are bit-identical to
    the eager calculator; a changed bond set falls back to the eager step and drops the graph."""
    lat, cart, z = synthetic.fcc_cu_supercell(2, 0.02, 8)
    model = _model(device)
    eager = M3GNetCalculator(model, 5.0, 4.0, skin=0.5, device=device)
    replay = M3GNetCalculator(model, 5.0, 4.0, skin=0.5, device=device, graph_replay=True, replay_after=2)
    rng = np.random.default_rng(2)
    kept = []
    for frame in range(12):
        e0, f0, s0 = eager.compute(lat, cart, z)
        e1, f1, s1 = replay.compute(lat, cart, z)
        assert torch.equal(e0, e1) and torch.equal(f0, f1) and torch.equal(s0, s1), frame
        kept.append(f1)                                   # returned tensors must survive later replays
        if frame == 8:
            cart = cart + rng.normal(0.0, 0.05, size=cart.shape)     # bond set changes: eager step, graph dropped
        else:
            cart = cart + rng.normal(0.0, 1e-5, size=cart.shape)     # far from any shell boundary: same bonds
    assert replay.n_replays >= 5, replay.n_replays
    e0, f0, _ = eager.compute(lat, cart, z)
    e1, f1, _ = replay.compute(lat, cart, z)
    assert torch.equal(e0, e1) and torch.equal(f0, f1)
    assert not torch.equal(kept[3], kept[4])              # distinct frames kept distinct values (clones, not views)
    opt = Fire(replay, lat, cart, z, dt=0.02, dt_max=0.05, max_move=0.01)
    ref = Fire(eager, lat, cart, z, dt=0.02, dt_max=0.05, max_move=0.01)
    for _ in range(10):
        opt.step()
        ref.step()
    assert torch.equal(opt.pos, ref.pos) and torch.equal(opt.forces, ref.forces)
