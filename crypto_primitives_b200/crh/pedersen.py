"""crh::pedersen::{Window, Parameters, CRH, TwoToOneCRH} and the injective-map compressor --
host mirror of R/crh/pedersen/mod.rs:23-209 and R/crh/injective_map/mod.rs:22-62 over the CUDA
library.  Byte inputs are numpy uint8 (n, len); points are affine (x, y) Montgomery limbs,
numpy uint64 (..., 2, 4)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as _f

import numpy as np

from .. import _native as N
from ..curves import JUBJUB, TECurve


@dataclass(frozen=True)
class Window:
    """pedersen::Window (mod.rs:23-26): WINDOW_SIZE, NUM_WINDOWS."""
    WINDOW_SIZE: int
    NUM_WINDOWS: int


def _u8(a):
    return a.ctypes.data_as(N.u8p)


def _u64(a):
    return a.ctypes.data_as(N.u64p)


class _Ctx:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                N.lib.cpb_pedersen_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass(eq=False)
class Parameters:
    """pedersen::Parameters{generators} (mod.rs:28-31); with `randomness_generator` it is
    commitment::pedersen::Parameters (R/commitment/pedersen/mod.rs:17-21).  generators:
    (NUM_WINDOWS, WINDOW_SIZE, 2, 4) affine Montgomery limbs; randomness_generator: (k, 2, 4)."""
    curve: TECurve
    window: Window
    generators: np.ndarray
    randomness_generator: np.ndarray | None = None
    chunk_bits: int = 0            # table lookup width on the device (0 = library default; 8..16); does not change results
    _ctx: dict = _f(default_factory=dict, repr=False)

    def __post_init__(self):
        self.generators = np.ascontiguousarray(self.generators, dtype=np.uint64)
        w = self.window
        # assert_eq!(parameters.generators.len(), W::NUM_WINDOWS), mod.rs:101-109
        assert self.generators.shape == (w.NUM_WINDOWS, w.WINDOW_SIZE, 2, 4), "Incorrect pp size for window params"
        if self.randomness_generator is not None:
            self.randomness_generator = np.ascontiguousarray(self.randomness_generator, dtype=np.uint64)
            assert self.randomness_generator.ndim == 3 and self.randomness_generator.shape[1:] == (2, 4)

    def context(self, device: int = 0):
        h = self._ctx.get(device)
        if h is None:
            out = N.vp()
            rg = self.randomness_generator
            N.check(N.lib.cpb_pedersen_ctx_create_ex(
                self.curve.id, self.window.WINDOW_SIZE, self.window.NUM_WINDOWS, _u64(self.generators),
                0 if rg is None else rg.shape[0], None if rg is None else _u64(rg), device, self.chunk_bits, C.byref(out)))
            h = _Ctx(out.value)
            self._ctx[device] = h
        return h.handle


def _points(curve: TECurve, pts) -> np.ndarray:
    return curve.base_field.elements([c for p in pts for c in p]).reshape(len(pts), 2, 4)


def generator_powers(curve: TECurve, num_powers: int, rng):
    """mod.rs:48-56: base = C::rand(rng); base, 2*base, 4*base, ..."""
    out, base = [], curve.random_point(rng)
    for _ in range(num_powers):
        out.append(base)
        base = curve.double(base)
    return out


def create_generators(curve: TECurve, window: Window, rng):
    """mod.rs:40-46."""
    return [generator_powers(curve, window.WINDOW_SIZE, rng) for _ in range(window.NUM_WINDOWS)]


def _map_len_error(e: N.CpbError, what: str):
    if e.status == N.CPB_BAD_LENGTH:
        # the reference panics here (mod.rs:82-89); Python callers get the same message as an exception
        raise ValueError(f"incorrect input length for {what}") from e
    raise e


class CRH:
    """CRHScheme{Input=[u8], Output=C::Affine, Parameters=Parameters<C>} (mod.rs:58-130)."""

    @staticmethod
    def setup(rng, window: Window, curve: TECurve = JUBJUB) -> Parameters:
        """mod.rs:64-74 (`rng`: any object with .field(q) -> int)."""
        gens = create_generators(curve, window, rng)
        return Parameters(curve, window, _points(curve, [p for w in gens for p in w]).reshape(window.NUM_WINDOWS, window.WINDOW_SIZE, 2, 4))

    @staticmethod
    def evaluate(parameters: Parameters, input, device: int = 0) -> np.ndarray:
        b = np.frombuffer(bytes(input), dtype=np.uint8)
        return CRH.evaluate_batch(parameters, b.reshape(1, -1), device)[0]

    @staticmethod
    def evaluate_batch(parameters: Parameters, inputs, device: int = 0) -> np.ndarray:
        """inputs (n, len) uint8 -> (n, 2, 4)."""
        inp = np.ascontiguousarray(inputs, dtype=np.uint8)
        assert inp.ndim == 2
        n, ln = inp.shape
        out = np.empty((n, 2, 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_pedersen_crh_batch(parameters.context(device), _u8(inp), ln, ln, _u64(out), n))
        except N.CpbError as e:
            _map_len_error(e, f"window params {parameters.window.WINDOW_SIZE}x{parameters.window.NUM_WINDOWS}")
        return out


class TwoToOneCRH:
    """TwoToOneCRHScheme{Input=[u8], Output=C::Affine} (mod.rs:132-198)."""

    setup = CRH.setup

    @staticmethod
    def evaluate(parameters: Parameters, left_input, right_input, device: int = 0) -> np.ndarray:
        """mod.rs:152-182: left || right copied into a zeroed (HALF+HALF)/8-byte buffer, then CRH."""
        l, r = bytes(left_input), bytes(right_input)
        assert len(l) == len(r), "left and right input should be of equal length"
        w = parameters.window
        half = (w.WINDOW_SIZE * w.NUM_WINDOWS) // 2
        buf = bytearray((half + half) // 8)
        data = (l + r)[:len(buf)]
        buf[:len(data)] = data
        return CRH.evaluate(parameters, bytes(buf), device)

    @staticmethod
    def compress(parameters: Parameters, left_input, right_input, device: int = 0) -> np.ndarray:
        """mod.rs:187-197: children serialised uncompressed on the device, then `evaluate`."""
        pair = np.stack([np.asarray(left_input, dtype=np.uint64).reshape(2, 4),
                         np.asarray(right_input, dtype=np.uint64).reshape(2, 4)])[None]
        return TwoToOneCRH.compress_batch(parameters, pair, device)[0]

    @staticmethod
    def compress_batch(parameters: Parameters, children, device: int = 0) -> np.ndarray:
        """children (n, 2, 2, 4) -> (n, 2, 4)."""
        ch = np.ascontiguousarray(children, dtype=np.uint64)
        assert ch.ndim == 4 and ch.shape[1:] == (2, 2, 4)
        out = np.empty((ch.shape[0], 2, 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_pedersen_two_to_one_batch(parameters.context(device), _u64(ch), _u64(out), ch.shape[0]))
        except N.CpbError as e:
            _map_len_error(e, "two-to-one window params")
        return out


class PedersenCRHCompressor:
    """PedersenCRHCompressor<C, TECompressor, W> (R/crh/injective_map/mod.rs:33-62): output = x-coordinate."""

    setup = CRH.setup

    @staticmethod
    def evaluate(parameters: Parameters, input, device: int = 0) -> np.ndarray:
        b = np.frombuffer(bytes(input), dtype=np.uint8)
        return PedersenCRHCompressor.evaluate_batch(parameters, b.reshape(1, -1), device)[0]

    @staticmethod
    def evaluate_batch(parameters: Parameters, inputs, device: int = 0) -> np.ndarray:
        inp = np.ascontiguousarray(inputs, dtype=np.uint8)
        n, ln = inp.shape
        out = np.empty((n, 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_pedersen_crh_x_batch(parameters.context(device), _u8(inp), ln, ln, _u64(out), n))
        except N.CpbError as e:
            _map_len_error(e, "window params")
        return out
