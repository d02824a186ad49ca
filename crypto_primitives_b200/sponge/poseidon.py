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


class PoseidonSponge:
    """Duplex sponge over the GPU permutation -- host mirror of PoseidonSponge<F> (R/sponge/poseidon/mod.rs:47-63):
    CryptographicSponge::{new, absorb, squeeze_bytes, squeeze_bits, squeeze_field_elements} (mod.rs:220-321) and
    FieldBasedCryptographicSponge::squeeze_native_field_elements (mod.rs:323-345), with the same mode bookkeeping
    (absorb_internal / squeeze_internal, mod.rs:124-186).  Only native field elements can be absorbed (the `Fp` and
    `&[Fp]` Absorb impls, R/sponge/absorb.rs:154-167, 284-292).  Every permutation is one GPU call of batch size 1:
    correct but slow -- a single transcript is sequential by nature; for many independent sponges use
    `absorb_squeeze_batch`."""

    def __init__(self, parameters: PoseidonConfig, device: int = 0):
        self.parameters = parameters
        self.device = device
        t = parameters.rate + parameters.capacity
        self.state = np.zeros((t, 4), dtype=np.uint64)
        self.mode = ("Absorbing", 0)            # DuplexSpongeMode::{Absorbing{next_absorb_index}, Squeezing{next_squeeze_index}}

    @classmethod
    def new(cls, parameters: PoseidonConfig, device: int = 0):
        return cls(parameters, device)

    # -- internals
    def _permute(self):
        out = np.empty_like(self.state)
        N.check(N.lib.cpb_poseidon_permute_batch(self.parameters.context(self.device), self.state.ctypes.data_as(N.u64p),
                                                 out.ctypes.data_as(N.u64p), 1))
        self.state = out

    def _add(self, lane: int, elem: np.ndarray):
        f = self.parameters.field
        a, b = f.to_ints(self.state[lane])[0], f.to_ints(elem)[0]
        self.state[lane] = f.elements([(a + b) % f.modulus])[0]

    def _absorb_internal(self, rate_start: int, elems: np.ndarray):           # mod.rs:124-153
        c = self.parameters
        rem = elems
        while True:
            if rate_start + len(rem) <= c.rate:
                for i, e in enumerate(rem):
                    self._add(c.capacity + i + rate_start, e)
                self.mode = ("Absorbing", rate_start + len(rem))
                return
            k = c.rate - rate_start
            for i, e in enumerate(rem[:k]):
                self._add(c.capacity + i + rate_start, e)
            self._permute()
            rem = rem[k:]
            rate_start = 0

    def _squeeze_internal(self, rate_start: int, n_out: int) -> np.ndarray:  # mod.rs:156-186
        c = self.parameters
        out = []
        rem = n_out
        while True:
            if rate_start + rem <= c.rate:
                out.extend(self.state[c.capacity + rate_start: c.capacity + rate_start + rem])
                self.mode = ("Squeezing", rate_start + rem)
                return np.array(out, dtype=np.uint64).reshape(-1, 4)
            k = c.rate - rate_start
            out.extend(self.state[c.capacity + rate_start: c.capacity + rate_start + k])
            rem -= k
            if rem != 0:
                self._permute()
            rate_start = 0

    # -- public surface
    def absorb(self, elems):
        """elems: (k, 4) Montgomery limbs (or a single (4,) element)."""
        e = np.asarray(elems, dtype=np.uint64).reshape(-1, 4)
        if e.shape[0] == 0:
            return
        kind, idx = self.mode
        if kind == "Absorbing":
            if idx == self.parameters.rate:
                self._permute()
                idx = 0
            self._absorb_internal(idx, e)
        else:
            self._absorb_internal(0, e)

    def squeeze_native_field_elements(self, num_elements: int) -> np.ndarray:
        kind, idx = self.mode
        if kind == "Absorbing":
            self._permute()
            return self._squeeze_internal(0, num_elements)
        if idx == self.parameters.rate:
            self._permute()
            idx = 0
        return self._squeeze_internal(idx, num_elements)

    squeeze_field_elements = squeeze_native_field_elements       # the native TypeId fast path, mod.rs:309-315

    def squeeze_bytes(self, num_bytes: int) -> bytes:
        """mod.rs:259-274."""
        f = self.parameters.field
        usable = (f.modulus_bit_size - 1) // 8
        n = (num_bytes + usable - 1) // usable
        out = b"".join(v.to_bytes(32, "little")[:usable] for v in f.to_ints(self.squeeze_native_field_elements(n)))
        return out[:num_bytes]

    def squeeze_bits(self, num_bits: int) -> list:
        """mod.rs:276-291."""
        f = self.parameters.field
        usable = f.modulus_bit_size - 1
        n = (num_bits + usable - 1) // usable
        bits = []
        for v in f.to_ints(self.squeeze_native_field_elements(n)):
            bits.extend(bool((v >> i) & 1) for i in range(usable))
        return bits[:num_bits]


def absorb_squeeze_batch(parameters: PoseidonConfig, inputs, num_squeeze: int, device: int = 0) -> np.ndarray:
    """n independent sponges in one kernel launch: new -> absorb(inputs[i]) -> squeeze_native_field_elements(num_squeeze).
    inputs (n, len, 4) -> (n, num_squeeze, 4)."""
    inp = np.ascontiguousarray(inputs, dtype=np.uint64)
    assert inp.ndim == 3 and inp.shape[2] == 4
    n, ln = inp.shape[0], inp.shape[1]
    out = np.empty((n, num_squeeze, 4), dtype=np.uint64)
    N.check(N.lib.cpb_poseidon_sponge_batch(parameters.context(device), inp.ctypes.data_as(N.u64p), ln, out.ctypes.data_as(N.u64p),
                                            num_squeeze, n))
    return out
