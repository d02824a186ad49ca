"""ctypes binding of the C restatement (oracle/cref/oracle_ref.c).  Test infrastructure only.

All field elements cross this boundary as numpy uint64 arrays of shape (..., 4):
little-endian limbs, Montgomery form (R = 2^256) -- the same interchange layout as the
product's C-ABI, so tests compare raw limbs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import fields as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle_ref.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cref", "oracle_ref.c")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_ref/liboracle_ref.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oref_field_new.restype = C.c_void_p
        L.oref_field_new.argtypes = [u64p]
        L.oref_field_free.argtypes = [C.c_void_p]
        L.oref_to_mont.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        L.oref_from_mont.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        L.oref_fe_mul_batch.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t]
        L.oref_fe_sqr_batch.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t]
        L.oref_fe_addsub_batch.argtypes = [C.c_void_p, u64p, u64p, u64p, u64p, C.c_size_t]
        L.oref_poseidon_new.restype = C.c_void_p
        L.oref_poseidon_new.argtypes = [u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, u64p, u64p]
        L.oref_poseidon_free.argtypes = [C.c_void_p]
        L.oref_poseidon_permute.argtypes = [C.c_void_p, u64p]
        L.oref_poseidon_crh_batch.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t, C.c_int]
        L.oref_poseidon_compress_batch.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_int]
        L.oref_poseidon_merkle.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_int]
        L.oref_poseidon_merkle.restype = C.c_int
        L.oref_pedersen_new.restype = C.c_void_p
        L.oref_pedersen_new.argtypes = [u64p, u64p, u64p, C.c_int, C.c_int, u64p, C.c_size_t, u64p]
        L.oref_pedersen_free.argtypes = [C.c_void_p]
        L.oref_pedersen_batch.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_size_t, u8p, u64p, C.c_size_t, C.c_int]
        L.oref_pedersen_batch.restype = C.c_int
        L.oref_pedersen_compress_batch.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_int]
        L.oref_pedersen_compress_batch.restype = C.c_int
        L.oref_pedersen_merkle.argtypes = [C.c_void_p, C.c_void_p, u8p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_int]
        L.oref_pedersen_merkle.restype = C.c_int
        L.oref_mixed_merkle.argtypes = [C.c_void_p, C.c_void_p, u8p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_int]
        L.oref_mixed_merkle.restype = C.c_int
        L.oref_bowe_hopwood_batch.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_size_t, u64p, C.c_size_t, C.c_int]
        L.oref_bowe_hopwood_batch.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _b(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


def limbs(x: int) -> np.ndarray:
    return np.array(F.to_limbs(x), dtype=np.uint64)


def ints_to_mont(vals, p: int) -> np.ndarray:
    """Python ints -> (n,4) uint64 Montgomery limbs (pure Python conversion)."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = F.to_limbs(F.to_mont(v % p, p))
    return out


def mont_to_ints(arr: np.ndarray, p: int) -> list[int]:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [F.from_mont(F.from_limbs(row), p) for row in a]


def field_ops(p: int, a: np.ndarray, b: np.ndarray):
    """Element-wise (a*b/R, a*a/R, a+b, a-b) mod p on (n,4) limb arrays through the C oracle's field routines."""
    a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
    n = a.shape[0]
    fld = lib().oref_field_new(_p(limbs(p)))
    mul, sqr, add, sub = (np.empty_like(a) for _ in range(4))
    lib().oref_fe_mul_batch(fld, _p(mul), _p(a), _p(b), n)
    lib().oref_fe_sqr_batch(fld, _p(sqr), _p(a), n)
    lib().oref_fe_addsub_batch(fld, _p(add), _p(sub), _p(a), _p(b), n)
    lib().oref_field_free(fld)
    return mul, sqr, add, sub


def synth_field_mont(seed: int, n: int, p: int) -> np.ndarray:
    """SURVEY.md §8d synthetic elements: SplitMix64 stream, 4 outputs = LE limbs, reduced mod p,
    then converted to Montgomery form by the C oracle.  Vectorised SplitMix64 in numpy."""
    raw = splitmix64_stream(seed, 4 * n).reshape(n, 4)
    fld = lib().oref_field_new(_p(limbs(p)))
    out = np.empty_like(raw)
    lib().oref_to_mont(fld, _p(out), _p(np.ascontiguousarray(raw)), n)
    lib().oref_field_free(fld)
    return out


def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & F.MASK64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_bytes(seed: int, n: int) -> np.ndarray:
    return splitmix64_stream(seed, (n + 7) // 8).view(np.uint8)[:n].copy()


class Poseidon:
    """Handle on a C-oracle PoseidonConfig (R/sponge/poseidon/mod.rs:26-45)."""

    def __init__(self, cfg):
        self.cfg = cfg
        p = cfg.p
        ark = ints_to_mont([x for row in cfg.ark for x in row], p)
        mds = ints_to_mont([x for row in cfg.mds for x in row], p)
        self.h = lib().oref_poseidon_new(_p(limbs(p)), cfg.rate, cfg.capacity, cfg.full_rounds,
                                         cfg.partial_rounds, cfg.alpha, _p(ark), _p(mds))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            lib().oref_poseidon_free(self.h)
            self.h = None

    def permute(self, state: np.ndarray) -> np.ndarray:
        st = np.ascontiguousarray(state, dtype=np.uint64).copy()
        lib().oref_poseidon_permute(self.h, _p(st))
        return st

    def crh_batch(self, inputs: np.ndarray, threads: int = 1) -> np.ndarray:
        """inputs (n, L, 4) -> (n, 4)."""
        inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
        n, L = inputs.shape[0], inputs.shape[1]
        out = np.empty((n, 4), dtype=np.uint64)
        lib().oref_poseidon_crh_batch(self.h, _p(inputs), L, _p(out), n, threads)
        return out

    def compress_batch(self, pairs: np.ndarray, threads: int = 1) -> np.ndarray:
        """pairs (n, 2, 4) -> (n, 4)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint64)
        n = pairs.shape[0]
        out = np.empty((n, 4), dtype=np.uint64)
        lib().oref_poseidon_compress_batch(self.h, _p(pairs), _p(out), n, threads)
        return out


def poseidon_merkle(leaf: Poseidon, node: Poseidon, leaves: np.ndarray, threads: int = 1):
    """leaves (n, L, 4) -> (leaf_nodes (n,4), non_leaf_nodes (n-1,4) heap order)."""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
    n, L = leaves.shape[0], leaves.shape[1]
    ln = np.empty((n, 4), dtype=np.uint64)
    nn = np.empty((n - 1, 4), dtype=np.uint64)
    rc = lib().oref_poseidon_merkle(leaf.h, node.h, _p(leaves), L, n, _p(ln), _p(nn), threads)
    if rc:
        raise ValueError("leaves.len() should be power of two and greater than one")
    return ln, nn


class Pedersen:
    """Handle on C-oracle Pedersen parameters over Jubjub (R/crh/pedersen/mod.rs:28-31,
    R/commitment/pedersen/mod.rs:17-21)."""

    def __init__(self, params, window):
        from . import jubjub as jj
        self.window = window
        q = jj.Q
        gens = [pt for w in params.generators for pt in w]
        gxy = ints_to_mont([c for pt in gens for c in pt], q)
        rnd = params.randomness_generator or []
        rxy = ints_to_mont([c for pt in rnd for c in pt], q) if rnd else np.zeros((1, 4), dtype=np.uint64)
        a = ints_to_mont([jj.A], q)
        d = ints_to_mont([jj.D], q)
        self.h = lib().oref_pedersen_new(_p(limbs(q)), _p(a), _p(d), window.window_size, window.num_windows,
                                         _p(gxy), len(rnd), _p(rxy))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oref_pedersen_free(self.h)
            self.h = None

    def batch(self, inputs: np.ndarray, randomness: np.ndarray | None = None, threads: int = 1) -> np.ndarray:
        """inputs (n, len) uint8; randomness (n, 32) uint8 LE scalars or None -> (n, 2, 4)."""
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        n, ln = inputs.shape
        out = np.empty((n, 2, 4), dtype=np.uint64)
        r = None
        if randomness is not None:
            randomness = np.ascontiguousarray(randomness, dtype=np.uint8)
            r = _b(randomness)
        rc = lib().oref_pedersen_batch(self.h, _b(inputs), ln, ln, r, _p(out), n, threads)
        if rc:
            raise ValueError("incorrect input length")
        return out

    def bowe_hopwood_batch(self, inputs: np.ndarray, threads: int = 1) -> np.ndarray:
        """Bowe-Hopwood CRH over these generators (window_size = chunks per segment): (n, len) uint8 -> (n, 4) x-coordinates."""
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        n, ln = inputs.shape
        out = np.empty((n, 4), dtype=np.uint64)
        if lib().oref_bowe_hopwood_batch(self.h, _b(inputs), ln, ln, _p(out), n, threads):
            raise ValueError("incorrect input bitlength")
        return out

    def compress_batch(self, children: np.ndarray, threads: int = 1) -> np.ndarray:
        """children (n, 2, 2, 4) affine point pairs -> (n, 2, 4)."""
        children = np.ascontiguousarray(children, dtype=np.uint64)
        n = children.shape[0]
        out = np.empty((n, 2, 4), dtype=np.uint64)
        rc = lib().oref_pedersen_compress_batch(self.h, _p(children), _p(out), n, threads)
        if rc:
            raise ValueError("incorrect input length")
        return out


def pedersen_merkle(leaf: Pedersen, node: Pedersen, leaves: np.ndarray, threads: int = 1):
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8)
    n, ln = leaves.shape
    lnodes = np.empty((n, 2, 4), dtype=np.uint64)
    nn = np.empty((n - 1, 2, 4), dtype=np.uint64)
    rc = lib().oref_pedersen_merkle(leaf.h, node.h, _b(leaves), ln, n, _p(lnodes), _p(nn), threads)
    if rc:
        raise ValueError("merkle build failed rc=%d" % rc)
    return lnodes, nn


def mixed_merkle(leaf: Pedersen, node: Poseidon, leaves: np.ndarray, threads: int = 1):
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8)
    n, ln = leaves.shape
    lnodes = np.empty((n, 4), dtype=np.uint64)
    nn = np.empty((n - 1, 4), dtype=np.uint64)
    rc = lib().oref_mixed_merkle(leaf.h, node.h, _b(leaves), ln, n, _p(lnodes), _p(nn), threads)
    if rc:
        raise ValueError("merkle build failed rc=%d" % rc)
    return lnodes, nn
