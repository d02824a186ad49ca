"""Bowe-Hopwood Pedersen CRH oracle (Python ints; test infrastructure only).  Restates
R/crh/bowe_hopwood/mod.rs:45-241 over `oracle.jubjub`.  PARITY UNPINNED (the reference holds only a
smoke test, mod.rs:253-271)."""
from __future__ import annotations

from . import jubjub as jj
from .fields import SplitMix64
from .pedersen import Parameters, Window, bytes_to_bits, synthetic_base

CHUNK_SIZE = 3


def create_generators(w: Window, rng: SplitMix64):
    """mod.rs:45-61: per segment a random base, then base, 16*base, 256*base, ..."""
    gens = []
    for _ in range(w.num_windows):
        seg, base = [], synthetic_base(rng)
        for _ in range(w.window_size):
            seg.append(base)
            for _ in range(4):
                base = jj.double(base)
        gens.append(seg)
    return gens


def setup(w: Window, seed: int) -> Parameters:
    return Parameters(create_generators(w, SplitMix64(seed)))


def crh_evaluate(params: Parameters, w: Window, inp: bytes) -> int:
    """mod.rs:115-185 -> x-coordinate of the affine sum."""
    if len(inp) * 8 > w.window_size * w.num_windows * CHUNK_SIZE:
        raise ValueError("incorrect input bitlength")
    bits = bytes_to_bits(inp)
    if len(bits) % CHUNK_SIZE:
        bits += [0] * (CHUNK_SIZE - len(bits) % CHUNK_SIZE)
    assert len(params.generators) == w.num_windows and all(len(g) == w.window_size for g in params.generators)
    acc = jj.IDENTITY
    seg_bits = w.window_size * CHUNK_SIZE
    for si in range(w.num_windows):                          # chunks(WINDOW_SIZE*3).zip(generators)
        seg = bits[si * seg_bits:(si + 1) * seg_bits]
        if not seg:
            break
        for ci in range(len(seg) // CHUNK_SIZE):             # chunks(3).zip(segment_generators)
            c0, c1, c2 = seg[3 * ci:3 * ci + 3]
            g = params.generators[si][ci]
            enc = g
            if c0:
                enc = jj.add(enc, g)
            if c1:
                enc = jj.add(enc, jj.double(g))
            if c2:
                enc = jj.neg(enc)
            acc = jj.add(acc, enc)
    return acc[0]


def two_to_one_evaluate(params: Parameters, w: Window, left: bytes, right: bytes) -> int:
    """mod.rs:200-226: buffer of INPUT_SIZE_BITS/8 bytes with INPUT_SIZE_BITS = WINDOW_SIZE*NUM_WINDOWS (:69)."""
    assert len(left) == len(right)
    buf = bytearray((w.window_size * w.num_windows) // 8)
    data = (bytes(left) + bytes(right))[:len(buf)]
    buf[:len(data)] = data
    return crh_evaluate(params, w, bytes(buf))


def two_to_one_compress(params: Parameters, w: Window, left_x: int, right_x: int) -> int:
    """mod.rs:228-240: children are base-field elements, serialised uncompressed (32-byte LE canonical)."""
    return two_to_one_evaluate(params, w, left_x.to_bytes(32, "little"), right_x.to_bytes(32, "little"))
