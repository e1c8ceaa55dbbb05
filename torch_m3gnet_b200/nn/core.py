from __future__ import annotations

import torch


class GatedMLP(torch.nn.Module):
    """Parameter container with the reference's layout (nn/core.py:6-62): ``dense`` and ``gate`` Sequentials
    whose Linear layers sit at even indices (0, 2, 4 …) so that state_dict keys are identical.  The arithmetic
    dense(x) * gate(x) is fused into the parent layer's CUDA kernel (csrc/conv.cu, threebody.cu, readout.cu)."""

    def __init__(self, in_features: int, dimensions: list[int], is_output: bool = False, use_bias: bool = True,
                 device: torch.device | None = None):
        super().__init__()
        self.in_features = in_features
        self.dimensions = dimensions
        self.is_output = is_output
        self.use_bias = use_bias
        self.dense = torch.nn.Sequential()
        self.gate = torch.nn.Sequential()
        widths = [in_features] + list(dimensions)
        last = len(dimensions) - 1
        for i in range(len(dimensions)):
            # creation order (dense then gate, layer by layer) matches the reference so that a given torch seed
            # yields the same initial weights
            self.dense.append(torch.nn.Linear(widths[i], widths[i + 1], bias=use_bias, device=device))
            if not (is_output and i == last):
                self.dense.append(torch.nn.SiLU())
            self.gate.append(torch.nn.Linear(widths[i], widths[i + 1], bias=use_bias, device=device))
            self.gate.append(torch.nn.Sigmoid() if i == last else torch.nn.SiLU())

    def linears(self, branch: str):
        seq = self.dense if branch == "dense" else self.gate
        return [m for m in seq if isinstance(m, torch.nn.Linear)]

    def forward(self, input):  # pragma: no cover - the fused kernels own the arithmetic
        raise RuntimeError("GatedMLP is evaluated inside the fused CUDA kernels of its parent layer")
