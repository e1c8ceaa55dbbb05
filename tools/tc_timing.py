"""Dev tool: per-phase cycle breakdown of the saved-activation backward gated-MLP kernel, thread 0 of every CTA
(library built with -DM3G_TC_TIMING; the recompute variant carries its own marks: M3G_TC_BWD_VARIANT=2).

  M3G_EXTRA_NVCC_FLAGS=-DM3G_TC_TIMING python -m torch_m3gnet_b200.csrc.build -f && python tools/tc_timing.py
  (or keep the instrumented build beside the product library and point M3G_LIB_PATH at it)
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch_m3gnet_b200 as m3g  # noqa: E402
from torch_m3gnet_b200 import _lib  # noqa: E402

NAMES = ["loop+prefetch", "T0 loads (z2, g_up, h)", "T5 adjoint math + TMEM st", "T5 barrier", "issue G3", "g_h + stash loads",
         "wait G3", "-", "T6 dz1 + TMEM st", "T6 barrier", "issue G4", "g_z1 stores", "wait G4", "T7 g_e stores", "-", "-"]


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = m3g.build_model(**bench.HP, device=dev)
    batch = bench.c2_workload(0).build(dev)
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
    order = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13]
    for k in order:
        print(f"{NAMES[k]:>16s} {per_tile[k]:9.0f} cyc/tile {100 * per_tile[k] / max(tot, 1):5.1f}%")
    print(f"{'total':>16s} {tot:9.0f} cyc/tile")


if __name__ == "__main__":
    main()
