"""Pedersen CRH / two-to-one / commitment oracle (Python ints; test infrastructure only).

Restates R/crh/pedersen/mod.rs:38-209, R/commitment/pedersen/mod.rs:44-105,
R/crh/injective_map/mod.rs:22-62 over the Jubjub oracle of `oracle.jubjub`.
PARITY UNPINNED: the reference holds no golden vectors for this path.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import jubjub as jj
from .fields import JUBJUB_FR, SplitMix64


@dataclass
class Window:
    """crh/pedersen/mod.rs:23-26."""
    window_size: int
    num_windows: int


@dataclass
class Parameters:
    """crh/pedersen/mod.rs:28-31: generators[w][j]; commitment adds randomness_generator (:17-21)."""
    generators: list
    randomness_generator: list | None = None


def synthetic_base(rng: SplitMix64):
    """draw y until a point exists, even root, times cofactor 8."""
    while True:
        P = jj.point_from_y(rng.field(jj.Q))
        if P is None:
            continue
        P = jj.mul(jj.COFACTOR, P)
        if P != jj.IDENTITY:
            return P


def generator_powers(num_powers: int, rng: SplitMix64):
    """crh/pedersen/mod.rs:48-56: base, 2*base, 4*base, ... (base drawn from our own RNG)."""
    out = []
    base = synthetic_base(rng)
    for _ in range(num_powers):
        out.append(base)
        base = jj.double(base)
    return out


def setup(w: Window, seed: int, commitment: bool = False) -> Parameters:
    """crh/pedersen/mod.rs:64-74 and commitment/pedersen/mod.rs:44-60 (randomness generator first)."""
    rng = SplitMix64(seed)
    rnd = generator_powers(JUBJUB_FR.bit_length(), rng) if commitment else None
    gens = [generator_powers(w.window_size, rng) for _ in range(w.num_windows)]
    return Parameters(gens, rnd)


def bytes_to_bits(b: bytes) -> list[int]:
    """crh/pedersen/mod.rs:200-209: bit i of byte k -> index 8k+i."""
    return [(byte >> i) & 1 for byte in b for i in range(8)]


def crh_evaluate(params: Parameters, w: Window, inp: bytes):
    """crh/pedersen/mod.rs:76-129."""
    nbits = w.window_size * w.num_windows
    if len(inp) * 8 > nbits:                                   # :82-89 panic
        raise ValueError("incorrect input length")
    if len(inp) * 8 < nbits:                                   # :94-99 zero pad to floor(bits/8)
        inp = bytes(inp) + bytes(max(0, nbits // 8 - len(inp)))
    assert len(params.generators) == w.num_windows             # :101-109
    bits = bytes_to_bits(inp)
    acc = jj.IDENTITY
    for wi in range(w.num_windows):                            # chunks(WINDOW_SIZE).zip(generators)
        chunk = bits[wi * w.window_size:(wi + 1) * w.window_size]
        if not chunk:
            break
        for bit, base in zip(chunk, params.generators[wi]):
            if bit:
                acc = jj.add(acc, base)
    return acc


def two_to_one_evaluate(params: Parameters, w: Window, left: bytes, right: bytes):
    """crh/pedersen/mod.rs:152-182."""
    assert len(left) == len(right)
    half = (w.window_size * w.num_windows) // 2
    buf = bytearray((half + half) // 8)
    data = bytes(left) + bytes(right)
    n = min(len(buf), len(data))
    buf[:n] = data[:n]
    return crh_evaluate(params, w, bytes(buf))


def two_to_one_compress(params: Parameters, w: Window, left_pt, right_pt):
    """crh/pedersen/mod.rs:187-197 (children serialised uncompressed, R/macros.rs)."""
    return two_to_one_evaluate(params, w, jj.serialize_uncompressed(left_pt),
                               jj.serialize_uncompressed(right_pt))


def commit(params: Parameters, w: Window, inp: bytes, randomness: int):
    """commitment/pedersen/mod.rs:62-105."""
    nbits = w.window_size * w.num_windows
    if len(inp) > nbits:                                       # :69-71 (bytes vs bits, as written)
        raise ValueError("incorrect input length")
    if len(inp) * 8 < nbits:
        inp = bytes(inp) + bytes(max(0, nbits // 8 - len(inp)))
    acc = crh_evaluate(Parameters(params.generators), w, inp)
    r = randomness % JUBJUB_FR
    for k, power in enumerate(params.randomness_generator):   # BitIteratorLE zip generators :93-100
        if (r >> k) & 1:
            acc = jj.add(acc, power)
    return acc


def te_compress(P) -> int:
    """crh/injective_map/mod.rs:24-31: TECompressor = x-coordinate."""
    return P[0]
