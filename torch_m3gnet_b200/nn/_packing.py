"""Kernel-side weight layouts, rebuilt lazily whenever a parameter tensor is replaced or modified
(``load_state_dict`` bumps ``_version``).  Pure layout work (transpose / concatenate), done once per change."""
from __future__ import annotations

from typing import Callable, Dict, Iterable

import torch


class PackedWeights:
    def __init__(self, params: Callable[[], Iterable[torch.Tensor]], pack: Callable[[], Dict]):
        self._params = params
        self._pack = pack
        self._sig = None
        self._cache = None

    def get(self) -> Dict:
        sig = tuple((p.data_ptr(), p._version, p.device) for p in self._params())
        if sig != self._sig:
            with torch.no_grad():
                self._cache = self._pack()
            self._sig = sig
        return self._cache


def module_params(module: torch.nn.Module) -> Callable[[], Iterable[torch.Tensor]]:
    """``lambda: list(module.parameters())`` without walking the module tree on every call: the (submodule, name)
    slots are listed once and read each time, so a replaced Parameter object is still seen."""
    slots = None

    def fetch():
        nonlocal slots
        if slots is None:
            slots = [(m, n) for m in module.modules() for n in m._parameters]
        return [m._parameters[n] for m, n in slots if m._parameters[n] is not None]

    return fetch


def t_(w: torch.Tensor) -> torch.Tensor:
    """(out,in) → contiguous (in,out)."""
    return w.detach().t().contiguous()


def c_(w: torch.Tensor) -> torch.Tensor:
    return w.detach().contiguous()
