from __future__ import annotations

import torch


class GatedMLP(torch.nn.Module):
    """The reference's gated MLP (nn/core.py:6-62): ``dense`` and ``gate`` Sequentials whose Linear layers sit at
    even indices (0, 2, 4 …) so that state_dict keys are identical.  Inside M3GNetConv / ThreeBodyInteration /
    AtomWiseReadout the arithmetic dense(x) * gate(x) is fused into the parent layer's CUDA kernel (csrc/conv_tc.cu,
    threebody_moment.cu, readout.cu); ``forward`` evaluates it on its own."""

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

    def forward(self, input):
        """dense(input) * gate(input) (reference nn/core.py:61-62) through the generic linear / activation kernels.
        The parent layers never take this route (their fused kernels own the arithmetic); it serves a GatedMLP that
        is called on its own.  Input gradients only, like every kernel-backed layer of this package."""
        from torch_m3gnet_b200.nn._functions import ActivationFn, LinearFn, MulFn

        def run(seq):
            y = input
            for m in seq:
                if isinstance(m, torch.nn.Linear):
                    y = LinearFn.apply(y, m.weight, m.bias)
                elif isinstance(m, torch.nn.SiLU):
                    y = ActivationFn.apply(y, 0)
                elif isinstance(m, torch.nn.Sigmoid):
                    y = ActivationFn.apply(y, 1)
                else:  # pragma: no cover
                    raise TypeError(f"unexpected layer {type(m).__name__} in a GatedMLP branch")
            return y

        with torch.cuda.device(input.device):
            return MulFn.apply(run(self.dense), run(self.gate))
