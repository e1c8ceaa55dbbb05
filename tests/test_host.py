"""CPU: host-side logic — the C-ABI library loads and exports every declared symbol, the module tree has the
reference's state_dict layout, data objects follow the reference's collate rules.  No kernel is launched."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.util import golden, graph_dict, state_dict_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from torch_m3gnet_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    protos = _lib.parse_header()
    text = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(m3g_\w+)\s*\(", re.sub(r"/\*.*?\*/", " ", text, flags=re.S)))
    assert declared == set(protos), declared ^ set(protos)
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/m3gnet_b200.h but not exported"
    assert len(protos) >= 50
    cdll.m3g_abi_version.restype = ctypes.c_int
    assert cdll.m3g_abi_version() == 1
    # pure host helpers can be called without a GPU: size of the saved-activation buffer (1 KB per bond, whole tiles)
    cdll.m3g_conv_tc_save_floats.restype = ctypes.c_int64
    cdll.m3g_conv_tc_save_floats.argtypes = [ctypes.c_int64]
    assert cdll.m3g_conv_tc_save_floats(0) == 0
    assert cdll.m3g_conv_tc_save_floats(1) == 128 * 256 and cdll.m3g_conv_tc_save_floats(129) == 2 * 128 * 256


def test_no_cpu_fallback():
    """The product path fails loudly on CPU tensors instead of falling back."""
    import torch_m3gnet_b200 as m3g

    g = golden("small_batch")
    gd = graph_dict(g)
    b = m3g.Batch(pos=gd["pos"], atom_types=gd["atom_types"], num_triplet_i=gd["num_triplet_i"],
                  edge_index=gd["edge_index"], edge_cell_shift=gd["edge_cell_shift"],
                  num_triplet_ij=gd["num_triplet_ij"], triplet_edge_index=gd["triplet_edge_index"],
                  lattice=gd["lattice"])
    b["batch"] = gd["batch"]
    rc = float(g["cutoff"])
    model = m3g.build_model(rc, rc, 2, 3, 93, 17, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(b)
    if not torch.cuda.is_available():
        from torch_m3gnet_b200.data.structure import Structure
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m3g.MaterialGraph.from_structure(Structure(np.eye(3) * 4, ["Cu"], [[0, 0, 0]]), 5.0, 4.0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "torch_m3gnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_state_dict_layout_and_seeded_init_match_reference():
    import torch_m3gnet_b200 as m3g

    g = golden("c1_default")
    ref_sd = state_dict_of(g)
    torch.manual_seed(0)
    model = m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3)
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys()) or set(sd.keys()) == set(ref_sd.keys())
    assert len(sd) == 80 and sum(v.numel() for v in sd.values()) == 227549  # docs/architecture.md:50
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref_sd[k].shape), k
        assert torch.equal(v, ref_sd[k]), f"{k}: same seed must give the reference's initial weights"
    assert len(list(model.named_buffers())) == 0  # constants are plain attributes
    model.load_state_dict(ref_sd)
    # the (noise-valued) Bessel normalisation table is reproduced on the host bit for bit (quirk Q1)
    assert np.array_equal(model.model[6].nsb.factors.numpy(), g["factors"])
    small = golden("small_batch")
    rc = float(small["cutoff"])
    m2 = m3g.build_model(rc, rc, 2, 3, 93, 17, 2)
    m2.load_state_dict(state_dict_of(small))
    assert np.array_equal(m2.model[6].nsb.factors.numpy(), small["factors"])


def test_build_model_argument_errors():
    import torch_m3gnet_b200 as m3g

    with pytest.raises(ValueError):
        m3g.build_model(5.0, 4.0, 10, 3, 95, 64, 1)  # l_max + 1 > 10 (nn/interaction.py:250-251)
    with pytest.raises(ValueError):
        m3g.build_model(5.0, 4.0, 3, 11, 95, 64, 1)  # n_max > 10 (:252-253)


def test_radial_constants_match_reference_known_answers():
    from torch_m3gnet_b200.nn.featurizer import radial_constants

    g = golden("basis")
    consts, em, dm, coeff = radial_constants(3, 5.0)
    assert np.array_equal(em.numpy(), g["em"]) and np.array_equal(dm.numpy(), g["dm"])
    assert np.array_equal(coeff.numpy(), g["coeff"])
    assert consts.shape == (4 + 3 * 3,)


def test_bessel_zero_table_matches_reference_after_float32_rounding():
    from torch_m3gnet_b200.nn.interaction import SPHERICAL_BESSEL_ZEROS

    g = golden("basis")
    ours = np.array(SPHERICAL_BESSEL_ZEROS)
    assert ours.shape == (10, 10)
    assert np.array_equal(ours.astype(np.float32), g["zeros"].astype(np.float32))
    assert np.abs(ours - g["zeros"]).max() < 1e-11


def test_collate_rules():
    """Batch.from_data_list follows data/material_graph.py:109-130 (index offsets, stacked lattice, batch)."""
    import torch_m3gnet_b200 as m3g
    from torch_m3gnet_b200.data import MaterialGraphKey as K

    def graph(n, e, t, seed):
        g0 = torch.Generator().manual_seed(seed)
        src = torch.sort(torch.randint(0, n, (e,), generator=g0))[0]
        return m3g.MaterialGraph(
            pos=torch.rand(n, 3, generator=g0), atom_types=torch.randint(0, 90, (n,), generator=g0),
            num_triplet_i=torch.zeros(n, dtype=torch.long), edge_index=torch.stack([src, torch.randint(0, n, (e,), generator=g0)]),
            edge_cell_shift=torch.zeros(e, 3, dtype=torch.int32), num_triplet_ij=torch.zeros(e, dtype=torch.int32),
            triplet_edge_index=torch.randint(0, e, (2, t), generator=g0), lattice=torch.eye(3) * (seed + 1))

    a, b = graph(4, 10, 7, 1), graph(2, 5, 3, 2)
    batch = m3g.Batch.from_data_list([a, b])
    assert batch.batch.tolist() == [0, 0, 0, 0, 1, 1] and batch.num_nodes == 6 and batch[K.NUM_EDGES] == 15
    assert batch[K.NUM_TRIPLETS] == 10 and batch.num_graphs == 2
    assert torch.equal(batch[K.EDGE_INDEX][:, 10:], b[K.EDGE_INDEX] + 4)
    assert torch.equal(batch[K.TRIPLET_EDGE_INDEX][:, 7:], b[K.TRIPLET_EDGE_INDEX] + 10)
    assert batch[K.LATTICE].shape == (2, 3, 3) and torch.equal(batch[K.LATTICE][1], torch.eye(3) * 3)
    assert batch.pos.shape == (6, 3)
    c = batch.clone()
    c[K.POS][0, 0] += 1.0
    assert not torch.equal(c[K.POS], batch[K.POS])
    assert batch.to(torch.device("cpu"))[K.POS].shape == (6, 3)
    # MaterialGraphKey names are the reference's
    assert (K.NODE_FEATURES, K.EDGE_ATTR, K.TOTAL_ENERGY, K.FORCES, K.BATCH) == ("x", "edge_attr", "total_energy", "forces", "batch")


def test_plan_signature_tracks_in_place_edits():
    from torch_m3gnet_b200.data.material_graph import GraphPlan
    import torch_m3gnet_b200 as m3g

    gd = graph_dict(golden("small_batch"))
    b = m3g.Batch(pos=gd["pos"], atom_types=gd["atom_types"], num_triplet_i=gd["num_triplet_i"],
                  edge_index=gd["edge_index"], edge_cell_shift=gd["edge_cell_shift"],
                  num_triplet_ij=gd["num_triplet_ij"], triplet_edge_index=gd["triplet_edge_index"].clone(),
                  lattice=gd["lattice"])
    b["batch"] = gd["batch"]
    s1 = GraphPlan.signature_of(b)
    b["triplet_edge_index"][0] = b["triplet_edge_index"][0].flip(0)  # in-place row assignment (tests/test_model.py:26-34)
    assert GraphPlan.signature_of(b) != s1


def test_structure_record():
    from torch_m3gnet_b200.data.structure import Structure

    s = Structure(np.eye(3) * 4.0, ["Ti", "O", 8], [[0, 0, 0], [0.5, 0.5, 0.5], [0.25, 0, 0]])
    assert len(s) == 3 and [site.specie.Z for site in s] == [22, 8, 8]
    assert np.allclose(s.cart_coords[1], [2, 2, 2]) and np.allclose(s.frac_coords[2], [0.25, 0, 0])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path: oracle/_ref when built, else the oracle port; no GPU needed) prints one JSON line with the
    keys the driver reads; the product arm's keys are checked on the GPU box by the bench itself."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "energy+forces atom-steps/sec" and d["unit"] == "atom-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C2")


def test_verlet_list_and_calculator_host_checks(monkeypatch):
    """Argument errors of the trajectory-side callers, and no CPU fallback for them either."""
    import torch_m3gnet_b200 as m3g
    from torch_m3gnet_b200 import calculator

    lat = np.eye(3)[None] * 4.0
    with pytest.raises(ValueError, match="Three body cutoff"):
        m3g.VerletList(lat, [29], [1], 4.0, 5.0)
    with pytest.raises(ValueError, match="skin"):
        m3g.VerletList(lat, [29], [1], 5.0, 4.0, skin=0.0)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m3g.VerletList(lat, [29], [1], 5.0, 4.0)
        calc = m3g.M3GNetCalculator(m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3), round_allocations=False)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            calc.compute(lat, np.zeros((1, 3)), [29])
    # the allocator policy of the user wins
    monkeypatch.setenv("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    assert calculator.stabilise_allocator() is False
    assert abs(calculator.ACC_UNIT - 1.602176634e-19 * 1e20 / (1.66053906660e-27 * 1e30)) < 1e-12


def test_packed_weights_follow_parameter_changes():
    """Kernel-side weight layouts are rebuilt when a parameter is modified in place, loaded, moved or replaced."""
    from torch_m3gnet_b200.nn._packing import PackedWeights, module_params, t_

    m = torch.nn.Sequential(torch.nn.Linear(3, 4, bias=False), torch.nn.Linear(4, 2))
    calls = []

    def pack():
        calls.append(1)
        return {"W0t": t_(m[0].weight), "W1t": t_(m[1].weight)}

    pw = PackedWeights(module_params(m), pack)
    a = pw.get()
    assert pw.get() is a and len(calls) == 1                       # cached
    with torch.no_grad():
        m[0].weight.mul_(2.0)                                      # in-place edit bumps _version
    b = pw.get()
    assert len(calls) == 2 and torch.equal(b["W0t"], m[0].weight.t())
    m.load_state_dict({k: v + 1 for k, v in m.state_dict().items()})
    c = pw.get()
    assert len(calls) == 3 and torch.equal(c["W1t"], m[1].weight.t())
    m[1].weight = torch.nn.Parameter(torch.zeros(2, 4))            # replaced Parameter object: the slot is re-read
    d = pw.get()
    assert len(calls) == 4 and torch.equal(d["W1t"], torch.zeros(4, 2))
    assert len(module_params(m)()) == 3


def test_bounded_sincos_restatement_accuracy():
    """csrc/threebody.cu::sincos_bounded (the radial three-body tables evaluate nine sin / cos pairs per member bond with
    it instead of sincosf) restated in numpy float32 with fused multiply-adds emulated in float64: against float64
    sin / cos over the argument range of the basis, x = z_ln r / r_c <= 12.4 (here [0, 16]), it stays within 1.5 ulp."""
    f = np.float32

    def fma(a, b, c):
        return f(np.float64(a) * np.float64(b) + np.float64(c))

    x = np.linspace(1e-6, 16.0, 2_000_001).astype(f)
    j = np.rint(f(x * f(0.636619747)))
    a = fma(j, f(-1.57079601e+00), x)
    a = fma(j, f(-3.13916473e-07), a)
    a = fma(j, f(-5.39030253e-15), a)
    s = f(a * a)
    r = f(2.86567956e-6)
    for coeff in (-1.98559923e-4, 8.33338592e-3, -1.66666672e-1):
        r = fma(r, s, f(coeff))
    sv = fma(r, f(a * s), a)
    c = f(2.44677067e-5)
    for coeff in (-1.38877297e-3, 4.16666567e-2, -5.00000000e-1, 1.0):
        c = fma(c, s, f(coeff))
    q = j.astype(np.int64)
    s0 = np.where(q & 1, c, sv)
    c0 = np.where(q & 1, sv, c)
    sin_v = np.where(q & 2, -s0, s0).astype(f)
    cos_v = np.where((q + 1) & 2, -c0, c0).astype(f)
    ref_s, ref_c = np.sin(x.astype(np.float64)), np.cos(x.astype(np.float64))
    ulp = lambda v: np.spacing(np.abs(v).astype(f)).astype(np.float64)  # noqa: E731
    assert (np.abs(sin_v - ref_s) / ulp(ref_s)).max() < 1.5
    assert (np.abs(cos_v - ref_c) / ulp(ref_c)).max() < 1.5
