"""Dev tool: latency of one energy + forces call on BASELINE config 1 (32-atom FCC Cu), eager vs CUDA-graph replay."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_m3gnet_b200 as m3g  # noqa: E402
from torch_m3gnet_b200 import synthetic  # noqa: E402
from torch_m3gnet_b200.graphed import GraphedStep  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=dev)
for reps in (2, 3, 5):
    lat, cart, z = synthetic.fcc_cu_supercell(reps, 0.05, 0)
    b = m3g.Batch.from_arrays(lat[None], cart, z, [len(cart)], 5.0, 4.0, device=dev)
    for _ in range(5):
        model(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        out = model(b)
    torch.cuda.synchronize()
    eager = (time.perf_counter() - t0) / n
    e_ref, f_ref = out["total_energy"].clone(), out["forces"].clone()
    step = GraphedStep(model, b)
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = step()
    torch.cuda.synchronize()
    graphed = (time.perf_counter() - t0) / n
    dE = (out["total_energy"] - e_ref).abs().max().item()
    dF = (out["forces"] - f_ref).abs().max().item()
    print(f"{len(cart):5d} atoms: eager {eager * 1e3:7.3f} ms  graph replay {graphed * 1e3:7.3f} ms  "
          f"({len(cart) / graphed:10.0f} atom-steps/s)  |dE|={dE:.1e} |dF|={dF:.1e}")
