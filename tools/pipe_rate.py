"""Dev tool: issue interval per SM sub-partition of the instruction kinds that pace the three-body MLP kernels
(FFMA with register / constant / immediate operands, packed FFMA2, MUFU, mma.sync tf32) — m3g_debug_pipe_rate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_m3gnet_b200 import _lib  # noqa: E402

KINDS = ["ffma 3 regs", "ffma shared multiplicands", "ffma constant-bank operand", "ffma2 3 reg pairs",
         "ffma2 shared multiplicands", "mufu ex2", "mma.sync m16n8k8 tf32", "ffma immediate"]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
iters = 2000
for kind, name in enumerate(KINDS):
    row = []
    for threads in (128, 256, 512, 1024):
        _lib.call("debug_pipe_rate", kind, threads, iters, out)
        torch.cuda.synchronize()
        warps_per_smsp = threads // 128
        cyc = out.item() / (iters * 32 * warps_per_smsp)
        row.append(f"{warps_per_smsp}w/SMSP {cyc:5.2f}")
    print(f"{name:30s} cycles per warp instruction per SMSP: " + "  ".join(row))
