"""`Absorb` encodings -- host mirror of R/sponge/absorb.rs:15-345 (R = /root/reference/crypto-primitives/src): how an
absorbable value becomes bytes (`to_sponge_bytes`) or native field elements (`to_sponge_field_elements`) before it
enters a sponge.  Pure host logic; the permutations run on the GPU (sponge/poseidon.py).

Python has no integer widths, so the Rust types are spelled out: `UInt(v, bits)`, `SInt(v, bits)`, `usize(v)`;
a bare `int` is Rust's default `i32`; `bytes` is `&[u8]`/`Vec<u8>`; a uint64 array (..., 4) is native field elements in
Montgomery limbs (wrap as `Elems(field, limbs)` to make the field explicit); `Point(curve, xy)` is a twisted-Edwards affine point; `None`/`Some(x)` is `Option`; `WithLength(x)`
selects the `AbsorbWithLength` form.  The [u8] -> field-element chunking ((MODULUS_BIT_SIZE-1)/8 bytes, little-endian)
and `Fp::serialize_compressed` (ceil(bits/8) LE bytes) are ark-ff 0.4 conventions.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..fields import Field


@dataclass(frozen=True)
class UInt:
    value: int
    bits: int


@dataclass(frozen=True)
class SInt:
    value: int
    bits: int


def usize(v: int) -> UInt:
    return UInt(v, 64)                       # absorb.rs:212-220


@dataclass(frozen=True, eq=False)
class Elems:
    field: Field
    limbs: np.ndarray                        # (k, 4) Montgomery limbs


@dataclass(frozen=True, eq=False)
class Point:
    """TEAffine<P> (absorb.rs:243-261): curve + (2, 4) affine Montgomery limbs of its base field."""
    curve: object
    xy: np.ndarray


@dataclass(frozen=True)
class Some:
    item: object


@dataclass(frozen=True)
class WithLength:
    item: object


def _norm(x):
    return SInt(x, 32) if isinstance(x, int) and not isinstance(x, bool) else x


def _len(x) -> int:
    return x.limbs.reshape(-1, 4).shape[0] if isinstance(x, Elems) else len(x)


def to_sponge_bytes(x) -> bytes:
    """Absorb::to_sponge_bytes."""
    x = _norm(x)
    if isinstance(x, bool):
        return bytes([int(x)])
    if isinstance(x, (UInt, SInt)):
        return (x.value % (1 << x.bits)).to_bytes(x.bits // 8, "little")
    if isinstance(x, Elems):
        nb = (x.field.modulus_bit_size + 7) // 8
        return b"".join(v.to_bytes(nb, "little") for v in x.field.to_ints(x.limbs))
    if isinstance(x, Point):
        nb = 8 * ((x.curve.base_field.modulus_bit_size + 63) // 64)          # into_bigint().to_bytes_le(): whole limbs
        return b"".join(v.to_bytes(nb, "little") for v in x.curve.base_field.to_ints(x.xy))
    if isinstance(x, (bytes, bytearray)):
        return bytes(x)
    if isinstance(x, str):
        return to_sponge_bytes(usize(len(x.encode()))) + x.encode()
    if x is None:
        return b"\x00"
    if isinstance(x, Some):
        return b"\x01" + to_sponge_bytes(x.item)
    if isinstance(x, WithLength):
        return to_sponge_bytes(usize(_len(x.item))) + to_sponge_bytes(x.item)
    if isinstance(x, (list, tuple)):
        return b"".join(to_sponge_bytes(i) for i in x)
    raise TypeError(f"not absorbable: {type(x).__name__}")


def _ints(x, field: Field) -> list:
    """Canonical integers of Absorb::to_sponge_field_elements::<F>; Montgomery-limb blocks are passed through as arrays."""
    x = _norm(x)
    p = field.modulus
    if isinstance(x, bool):
        return [int(x)]
    if isinstance(x, (UInt, SInt)):
        return [x.value % p]                                                   # negative: -F::from(|v|), :195-201
    if isinstance(x, np.ndarray):
        return [np.asarray(x, dtype=np.uint64).reshape(-1, 4)]
    if isinstance(x, Elems):
        if x.field.modulus != p:
            raise ValueError("Trying to absorb non-native field elements.")    # field_cast(..).unwrap(), :106-122
        return [np.asarray(x.limbs, dtype=np.uint64).reshape(-1, 4)]
    if isinstance(x, Point):
        if x.curve.base_field.modulus != p:
            raise ValueError("Trying to absorb non-native field elements.")
        return [np.asarray(x.xy, dtype=np.uint64).reshape(2, 4)]               # [x, y], :258-260
    if isinstance(x, (bytes, bytearray)):
        b = len(x).to_bytes(8, "little") + bytes(x)                            # :137-141
        k = (p.bit_length() - 1) // 8
        return [int.from_bytes(b[i:i + k], "little") for i in range(0, len(b), k)]
    if isinstance(x, str):
        return _ints(x.encode(), field)
    if x is None:
        return [0]
    if isinstance(x, Some):
        return [1] + _ints(x.item, field)
    if isinstance(x, WithLength):
        return _ints(usize(_len(x.item)), field) + _ints(x.item, field)
    if isinstance(x, (list, tuple)):
        if len(x) and all(isinstance(i, UInt) and i.bits == 8 for i in x):
            # &[u8] / Vec<u8> take the u8 batch specialisation (absorb.rs:137-141): length prefix + packed chunks
            return _ints(bytes(i.value % 256 for i in x), field)
        out = []
        for i in x:
            out += _ints(i, field)
        return out
    raise TypeError(f"not absorbable: {type(x).__name__}")


def to_sponge_field_elements(x, field: Field) -> np.ndarray:
    """Absorb::to_sponge_field_elements::<F> as (k, 4) Montgomery limbs of `field`."""
    parts, run = [], []
    for item in _ints(x, field):
        if isinstance(item, np.ndarray):
            if run:
                parts.append(field.elements(run))
                run = []
            parts.append(item)
        else:
            run.append(item)
    if run:
        parts.append(field.elements(run))
    return np.concatenate(parts) if parts else np.zeros((0, 4), dtype=np.uint64)
