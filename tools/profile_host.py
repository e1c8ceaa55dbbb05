"""Host-side cost of one small-system step (32-atom cell): wall time per calculator / model call and a cProfile of
the Python path (ctypes launches, autograd bookkeeping).  python tools/profile_host.py"""
import cProfile, pstats, sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_m3gnet_b200 as m3g
from torch_m3gnet_b200 import synthetic
dev = torch.device("cuda:0")
lat, cart, z = synthetic.fcc_cu_supercell(2, 0.05, 4)
model = m3g.build_model(5.0, 4.0, 3, 3, 95, 64, 3, device=dev)
calc = m3g.M3GNetCalculator(model, device=dev)
for _ in range(20): calc.compute(lat, cart, z)
torch.cuda.synchronize()
import time
t=time.time()
for _ in range(200): calc.compute(lat, cart, z)
torch.cuda.synchronize(); print("ms/step", (time.time()-t)/200*1e3)
b = calc.neighbor_list.update(cart)
t=time.time()
for _ in range(200): model(b)
torch.cuda.synchronize(); print("model only ms", (time.time()-t)/200*1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): calc.compute(lat, cart, z)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
