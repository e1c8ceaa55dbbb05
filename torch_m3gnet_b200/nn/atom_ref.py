from __future__ import annotations

import torch

from torch_m3gnet_b200._lib import call
from torch_m3gnet_b200.data import MaterialGraphKey as K
from torch_m3gnet_b200.data.material_graph import get_plan


class AtomRef(torch.nn.Module):
    """Elemental reference energies E_elem[Z-1] (reference nn/atom_ref.py:10-29).  As in the reference the
    table is a plain attribute, not a buffer (it is not part of the state_dict)."""

    def __init__(self, elemental_energies: torch.Tensor, device: torch.device | None = None):
        super().__init__()
        self.elemental_energies = elemental_energies.to(device)
        self._table_cache = None

    def forward(self, graph):
        plan = get_plan(graph)
        src = self.elemental_energies
        sig = (src.data_ptr(), src._version, plan.device)
        if self._table_cache is None or self._table_cache[0] != sig:
            self._table_cache = (sig, src.detach().to(device=plan.device, dtype=torch.float32).contiguous())
        table = self._table_cache[1]
        plan.check_types(int(table.numel()), "the elemental-energy table")
        out = torch.empty(plan.N, dtype=torch.float32, device=plan.device)
        call("atomref_fwd", table, plan.types, plan.N, out)
        graph[K.ELEMENTAL_ENERGIES] = out
        return graph
