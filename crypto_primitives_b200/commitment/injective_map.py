"""commitment::injective_map::PedersenCommCompressor -- host mirror of R/commitment/injective_map/mod.rs:11-44 over the CUDA
library: a Pedersen commitment followed by the injective map of R/crh/injective_map/mod.rs:23-31 (TECompressor: the affine
x-coordinate)."""
from __future__ import annotations

import numpy as np

from .pedersen import Commitment, Parameters


class PedersenCommCompressor:
    """PedersenCommCompressor<C, TECompressor, W> (R/commitment/injective_map/mod.rs:11-44): the commitment's
    x-coordinate (TECompressor::injective_map, R/crh/injective_map/mod.rs:23-31)."""

    setup = Commitment.setup

    @staticmethod
    def commit(parameters: Parameters, input, randomness: int, device: int = 0) -> np.ndarray:
        return Commitment.commit(parameters, input, randomness, device)[0]

    @staticmethod
    def commit_batch(parameters: Parameters, inputs, randomness_le32, device: int = 0) -> np.ndarray:
        """(n, len) uint8, (n, 32) uint8 -> (n, 4): x of each commitment."""
        return np.ascontiguousarray(Commitment.commit_batch(parameters, inputs, randomness_le32, device)[:, 0, :])
