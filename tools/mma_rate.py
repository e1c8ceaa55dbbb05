"""Dev tool: tcgen05.mma kind::tf32 issue / completion rate per instruction width (debug ABI entry)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_m3gnet_b200 import _lib  # noqa: E402

out = torch.zeros(2, dtype=torch.int64, device="cuda")
for a_tmem in (0, 1):
    for N in (64, 128, 256):
        for n in (64, 256):
            _lib.call("debug_mma_rate", N, a_tmem, n, out)
            torch.cuda.synchronize()
            i, c = out.tolist()
            print(f"a_tmem={a_tmem} N={N:3d} n_mma={n:3d}: issue {i / n:6.1f} cyc/mma, complete {c / n:6.1f} cyc/mma")

# cta_group::2: M = 256 over a CTA pair (each SM: its own 128 rows x N)
if "--pair" in sys.argv:
    for N in (64, 128, 256):
        for n in (64, 256):
            _lib.call("debug_mma_rate2", N, n, out)
            torch.cuda.synchronize()
            i, c = out.tolist()
            print(f"cta_group::2 M=256 N={N:3d} n_mma={n:3d}: issue {i / n:6.1f} cyc/mma, complete {c / n:6.1f} cyc/mma")
