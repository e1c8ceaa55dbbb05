"""Dev tool: per-phase cycle breakdown of the backward gated-MLP kernel (library built with -DM3G_TC_TIMING).

  M3G_EXTRA_NVCC_FLAGS=-DM3G_TC_TIMING python -m torch_m3gnet_b200.csrc.build -f && python tools/tc_timing.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch_m3gnet_b200 as m3g  # noqa: E402
from torch_m3gnet_b200 import _lib  # noqa: E402

NAMES = ["loop", "T1 e->TMEM", "issue G1", "wait G1", "T3 act", "issue G2", "wait G2", "T5 adjoint", "issue G3",
         "wait G3", "T6 dz1", "issue G4", "wait G4", "T7 out", "T2/T4 gathers", "pre-wait G3/G4"]


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = m3g.build_model(**bench.HP, device=dev)
    batch, _, _ = bench.build_inputs(dev, 0)
    for _ in range(2):
        model(batch)
    buf = (ctypes.c_int64 * 16)()
    cdll = _lib.LIB.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cdll.m3g_debug_tc_timing(buf, 1, stream)
    steps = 3
    for _ in range(steps):
        model(batch)
    cdll.m3g_debug_tc_timing(buf, 1, stream)
    vals = list(buf)
    n_tiles = (batch._plan.E + 127) // 128
    per_tile = [v / (steps * 6 * n_tiles) for v in vals]  # 6 backward launches per step
    tot = sum(per_tile)
    order = [0, 1, 2, 14, 3, 4, 5, 6, 7, 8, 15, 9, 10, 11, 12, 13]
    for k in order:
        print(f"{NAMES[k]:>16s} {per_tile[k]:9.0f} cyc/tile {100 * per_tile[k] / max(tot, 1):5.1f}%")
    print(f"{'total':>16s} {tot:9.0f} cyc/tile")


if __name__ == "__main__":
    main()
