"""PoseidonConfig and parameter generation -- host mirror of R/sponge/poseidon/mod.rs:26-45,189-217
and R/sponge/poseidon/traits.rs:59-146 (R = /root/reference/crypto-primitives/src).  The Grain LFSR
and Cauchy-matrix arithmetic run in the library's host code (csrc/poseidon_host.hpp); permutations
run only on the GPU (see crh/poseidon.py).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as _f

import numpy as np

from .. import _native as N
from ..fields import Field


@dataclass(eq=False)
class PoseidonConfig:
    """Same members as the reference struct (mod.rs:26-45); `ark` is (full+partial, t, 4) and `mds`
    (t, t, 4) uint64 Montgomery limbs; `field` replaces the Rust type parameter."""
    field: Field
    full_rounds: int
    partial_rounds: int
    alpha: int
    mds: np.ndarray
    ark: np.ndarray
    rate: int
    capacity: int
    _ctx: dict = _f(default_factory=dict, repr=False)

    def __post_init__(self):
        t = self.rate + self.capacity
        self.ark = np.ascontiguousarray(self.ark, dtype=np.uint64)
        self.mds = np.ascontiguousarray(self.mds, dtype=np.uint64)
        # PoseidonConfig::new asserts, mod.rs:198-206
        assert self.ark.shape == (self.full_rounds + self.partial_rounds, t, 4), "ark shape"
        assert self.mds.shape == (t, t, 4), "mds shape"

    @classmethod
    def new(cls, field, full_rounds, partial_rounds, alpha, mds, ark, rate, capacity):
        """Argument order of PoseidonConfig::new (mod.rs:189-197)."""
        return cls(field, full_rounds, partial_rounds, alpha, mds, ark, rate, capacity)

    @classmethod
    def from_ints(cls, field: Field, full_rounds, partial_rounds, alpha, mds, ark, rate, capacity):
        t = rate + capacity
        a = field.elements([x for row in ark for x in row]).reshape(full_rounds + partial_rounds, t, 4)
        m = field.elements([x for row in mds for x in row]).reshape(t, t, 4)
        return cls(field, full_rounds, partial_rounds, alpha, m, a, rate, capacity)

    # -- device context (created on first use, one per device)
    def context(self, device: int = 0):
        h = self._ctx.get(device)
        if h is None:
            out = N.vp()
            N.check(N.lib.cpb_poseidon_ctx_create(
                self.field.id, self.rate, self.capacity, self.full_rounds, self.partial_rounds, self.alpha,
                self.ark.ctypes.data_as(N.u64p), self.mds.ctypes.data_as(N.u64p), device, C.byref(out)))
            h = _Ctx(out.value)
            self._ctx[device] = h
        return h.handle


class _Ctx:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                N.lib.cpb_poseidon_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def find_poseidon_ark_and_mds(field: Field, prime_bits: int, rate: int, full_rounds: int, partial_rounds: int,
                              skip_matrices: int):
    """traits.rs:105-146 -> (ark (R, t, 4), mds (t, t, 4)) Montgomery limbs."""
    t = rate + 1
    ark = np.empty((full_rounds + partial_rounds, t, 4), dtype=np.uint64)
    mds = np.empty((t, t, 4), dtype=np.uint64)
    N.check(N.lib.cpb_poseidon_find_ark_and_mds(field.id, prime_bits, rate, full_rounds, partial_rounds, skip_matrices,
                                                ark.ctypes.data_as(N.u64p), mds.ctypes.data_as(N.u64p)))
    return ark, mds


def get_default_poseidon_parameters(field: Field, rate: int, optimized_for_weights: bool):
    """PoseidonDefaultConfigField::get_default_poseidon_parameters (traits.rs:59-103).  The entry
    tables are those of the reference's BLS12-381 Fr test field (R/sponge/test.rs:13-32); as in the
    reference, a rate without an entry yields None."""
    alpha = C.c_uint64()
    rf, rp, skip = C.c_int(), C.c_int(), C.c_int()
    st = N.lib.cpb_poseidon_default_entry(rate, int(bool(optimized_for_weights)), C.byref(alpha), C.byref(rf), C.byref(rp), C.byref(skip))
    if st != N.CPB_OK:
        return None
    ark, mds = find_poseidon_ark_and_mds(field, field.modulus_bit_size, rate, rf.value, rp.value, skip.value)
    return PoseidonConfig(field, rf.value, rp.value, alpha.value, mds, ark, rate, 1)
