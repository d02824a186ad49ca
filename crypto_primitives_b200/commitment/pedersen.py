"""commitment::pedersen::{Parameters, Randomness, Commitment} -- host mirror of
R/commitment/pedersen/mod.rs:17-106 over the CUDA library."""
from __future__ import annotations

import numpy as np

from .. import _native as N
from ..crh import pedersen as ped
from ..crh.pedersen import Parameters, Window  # noqa: F401  (same struct with randomness_generator set)
from ..curves import JUBJUB, TECurve


def randomness_bytes(curve: TECurve, values) -> np.ndarray:
    """Randomness<C>(pub C::ScalarField) -> the 32-byte little-endian canonical integers the ABI takes
    (`randomness.0.into_bigint()`, mod.rs:93)."""
    out = np.zeros((len(values), 32), dtype=np.uint8)
    for i, v in enumerate(values):
        out[i] = np.frombuffer((int(v) % curve.scalar_modulus).to_bytes(32, "little"), dtype=np.uint8)
    return out


class Commitment:
    """CommitmentScheme{Parameters, Randomness, Output=C::Affine} (mod.rs:38-106)."""

    @staticmethod
    def setup(rng, window: Window, curve: TECurve = JUBJUB) -> Parameters:
        """mod.rs:44-60: randomness generator (MODULUS_BIT_SIZE powers) first, then the window generators."""
        rnd = ped.generator_powers(curve, curve.scalar_modulus_bit_size, rng)
        gens = ped.create_generators(curve, window, rng)
        g = ped._points(curve, [p for w in gens for p in w]).reshape(window.NUM_WINDOWS, window.WINDOW_SIZE, 2, 4)
        return Parameters(curve, window, g, ped._points(curve, rnd))

    @staticmethod
    def commit(parameters: Parameters, input, randomness: int, device: int = 0) -> np.ndarray:
        b = np.frombuffer(bytes(input), dtype=np.uint8).reshape(1, -1)
        return Commitment.commit_batch(parameters, b, randomness_bytes(parameters.curve, [randomness]), device)[0]

    @staticmethod
    def commit_batch(parameters: Parameters, inputs, randomness_le32, device: int = 0) -> np.ndarray:
        """inputs (n, len) uint8, randomness_le32 (n, 32) uint8 -> (n, 2, 4)."""
        inp = np.ascontiguousarray(inputs, dtype=np.uint8)
        rnd = np.ascontiguousarray(randomness_le32, dtype=np.uint8)
        n, ln = inp.shape
        assert rnd.shape == (n, 32)
        out = np.empty((n, 2, 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_pedersen_commit_batch(parameters.context(device), inp.ctypes.data_as(N.u8p), ln, ln,
                                                    rnd.ctypes.data_as(N.u8p), out.ctypes.data_as(N.u64p), n))
        except N.CpbError as e:
            if e.status == N.CPB_BAD_LENGTH:
                raise ValueError(f"incorrect input length: {ln}") from e       # panic at mod.rs:69-71
            raise
        return out
