"""crh::bowe_hopwood::{Parameters, CRH, TwoToOneCRH} -- host mirror of R/crh/bowe_hopwood/mod.rs:31-241 over the
CUDA library.  Byte inputs numpy uint8 (n, len); outputs are x-coordinates, one base-field element (4 limbs)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as _f

import numpy as np

from .. import _native as N
from ..curves import JUBJUB, TECurve
from .pedersen import Window, _points, _u8, _u64

CHUNK_SIZE = 3


class _Ctx:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                N.lib.cpb_bowe_hopwood_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass(eq=False)
class Parameters:
    """bowe_hopwood::Parameters{generators} (mod.rs:33-37): (NUM_WINDOWS, WINDOW_SIZE, 2, 4) affine Montgomery limbs;
    WINDOW_SIZE = 3-bit chunks per segment."""
    curve: TECurve
    window: Window
    generators: np.ndarray
    _ctx: dict = _f(default_factory=dict, repr=False)

    def __post_init__(self):
        self.generators = np.ascontiguousarray(self.generators, dtype=np.uint64)
        w = self.window
        assert self.generators.shape == (w.NUM_WINDOWS, w.WINDOW_SIZE, 2, 4), "Incorrect pp size for window params"

    def context(self, device: int = 0):
        h = self._ctx.get(device)
        if h is None:
            out = N.vp()
            N.check(N.lib.cpb_bowe_hopwood_ctx_create(self.curve.id, self.window.WINDOW_SIZE, self.window.NUM_WINDOWS,
                                                      _u64(self.generators), device, C.byref(out)))
            h = _Ctx(out.value)
            self._ctx[device] = h
        return h.handle


def calculate_num_chunks_in_segment(scalar_modulus: int) -> int:
    """mod.rs:83-94: how many 3-bit chunks keep the encoded scalar below (r-1)/2."""
    upper, c, rng = (scalar_modulus - 1) // 2, 0, 2
    while rng < upper:
        rng <<= 4
        c += 1
    return c


def create_generators(curve: TECurve, window: Window, rng):
    """mod.rs:45-61: per segment a random base, then base, 16*base, 16^2*base, ..."""
    gens = []
    for _ in range(window.NUM_WINDOWS):
        seg, base = [], curve.random_point(rng)
        for _ in range(window.WINDOW_SIZE):
            seg.append(base)
            for _ in range(4):
                base = curve.double(base)
        gens.append(seg)
    return gens


class CRH:
    """CRHScheme{Input=[u8], Output=P::BaseField} (mod.rs:76-186)."""

    @staticmethod
    def setup(rng, window: Window, curve: TECurve = JUBJUB) -> Parameters:
        if window.WINDOW_SIZE > calculate_num_chunks_in_segment(curve.scalar_modulus):        # panic at mod.rs:96-103
            raise ValueError("Bowe-Hopwood-PedersenCRH hash must have a window size resulting in scalars < (p-1)/2")
        gens = create_generators(curve, window, rng)
        return Parameters(curve, window, _points(curve, [p for s in gens for p in s]).reshape(window.NUM_WINDOWS, window.WINDOW_SIZE, 2, 4))

    @staticmethod
    def evaluate(parameters: Parameters, input, device: int = 0) -> np.ndarray:
        b = np.frombuffer(bytes(input), dtype=np.uint8)
        return CRH.evaluate_batch(parameters, b.reshape(1, -1), device)[0]

    @staticmethod
    def evaluate_batch(parameters: Parameters, inputs, device: int = 0) -> np.ndarray:
        inp = np.ascontiguousarray(inputs, dtype=np.uint8)
        assert inp.ndim == 2
        n, ln = inp.shape
        out = np.empty((n, 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_bowe_hopwood_crh_batch(parameters.context(device), _u8(inp), ln, max(ln, 1), _u64(out), n))
        except N.CpbError as e:
            if e.status == N.CPB_BAD_LENGTH:
                raise ValueError(f"incorrect input bitlength {ln * 8}") from e          # panic at mod.rs:121-129
            raise
        return out


class TwoToOneCRH:
    """TwoToOneCRHScheme{Input=[u8], Output=P::BaseField} (mod.rs:188-241)."""

    setup = CRH.setup

    @staticmethod
    def evaluate(parameters: Parameters, left_input, right_input, device: int = 0) -> np.ndarray:
        l, r = bytes(left_input), bytes(right_input)
        assert len(l) == len(r), "left and right input should be of equal length"
        w = parameters.window
        buf = bytearray((w.WINDOW_SIZE * w.NUM_WINDOWS) // 8)        # INPUT_SIZE_BITS / 8, mod.rs:218
        data = (l + r)[:len(buf)]
        buf[:len(data)] = data
        return CRH.evaluate(parameters, bytes(buf), device)

    @staticmethod
    def compress(parameters: Parameters, left_input, right_input, device: int = 0) -> np.ndarray:
        pair = np.stack([np.asarray(left_input, dtype=np.uint64).reshape(4), np.asarray(right_input, dtype=np.uint64).reshape(4)])[None]
        return TwoToOneCRH.compress_batch(parameters, pair, device)[0]

    @staticmethod
    def compress_batch(parameters: Parameters, children, device: int = 0) -> np.ndarray:
        """children (n, 2, 4) -> (n, 4)."""
        ch = np.ascontiguousarray(children, dtype=np.uint64)
        assert ch.ndim == 3 and ch.shape[1:] == (2, 4)
        out = np.empty((ch.shape[0], 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_bowe_hopwood_two_to_one_batch(parameters.context(device), _u64(ch), _u64(out), ch.shape[0]))
        except N.CpbError as e:
            if e.status == N.CPB_BAD_LENGTH:
                raise ValueError("incorrect input bitlength") from e
            raise
        return out
