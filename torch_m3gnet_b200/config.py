from __future__ import annotations

from dataclasses import dataclass


@dataclass
class RunConfig:
    """Model hyper-parameters of the reference's RunConfig (config.py:10-17).  The training fields of the
    reference (epochs, learning rate, loss weights …) belong to its Lightning harness, which is out of scope."""

    root: str = ""
    cutoff: float = 5.0
    threebody_cutoff: float = 4.0
    l_max: int = 3
    n_max: int = 3
    num_types: int = 95
    embedding_dim: int = 64
    num_blocks: int = 3
