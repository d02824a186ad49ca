"""TEST INFRASTRUCTURE ONLY -- oracle for the `Absorb` encodings and the generic squeeze helpers of the sponge
interface (R = /root/reference/crypto-primitives/src): R/sponge/absorb.rs:15-345, R/sponge/mod.rs:30-153.

Values are Python ints (canonical field elements) and bytes.  Two conventions come from ark-ff 0.4 (dependency, not in
/root/reference) and are restated from its published behaviour -- parity for them is UNPINNED:
  * `ToConstraintField<F> for [u8]`: split into chunks of (MODULUS_BIT_SIZE-1)/8 bytes, each read as a little-endian integer;
  * `Fp::serialize_compressed`: ceil(MODULUS_BIT_SIZE/8) little-endian bytes of the canonical value.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class UInt:          # u8/u16/u32/u64/u128; usize is UInt(v, 64) (absorb.rs:212-220)
    value: int
    bits: int


@dataclass(frozen=True)
class SInt:          # i8..i128, isize = SInt(v, 64)
    value: int
    bits: int


@dataclass(frozen=True)
class Fe:            # a prime-field element: canonical value and its modulus
    value: int
    p: int


@dataclass(frozen=True)
class TEPoint:       # TEAffine<P> over a prime base field: canonical x, y and the base-field modulus (absorb.rs:243-261)
    x: int
    y: int
    q: int


@dataclass(frozen=True)
class Some:
    item: object


@dataclass(frozen=True)
class WithLength:    # AbsorbWithLength::to_sponge_*_with_length (absorb.rs:84-103) of a list / byte string
    item: object


def _norm(x):
    return SInt(x, 32) if isinstance(x, int) and not isinstance(x, bool) else x     # Rust's default integer type


def bytes_to_field_elements(b: bytes, p: int) -> list[int]:
    """ark-ff `ToConstraintField for [u8]` (restated)."""
    k = (p.bit_length() - 1) // 8
    return [int.from_bytes(b[i:i + k], "little") for i in range(0, len(b), k)]


def to_sponge_bytes(x) -> bytes:
    x = _norm(x)
    if isinstance(x, bool):
        return bytes([int(x)])                                                 # :145-147
    if isinstance(x, (UInt, SInt)):
        return (x.value % (1 << x.bits)).to_bytes(x.bits // 8, "little")       # :172-174, :191-193 (two's complement)
    if isinstance(x, Fe):
        return x.value.to_bytes((x.p.bit_length() + 7) // 8, "little")         # :155-157
    if isinstance(x, TEPoint):
        nb = 8 * ((x.q.bit_length() + 63) // 64)                               # BigInt::to_bytes_le: whole limbs
        return x.x.to_bytes(nb, "little") + x.y.to_bytes(nb, "little")         # :247-256
    if isinstance(x, (bytes, bytearray)):
        return bytes(x)                                                        # :133-135
    if isinstance(x, str):
        return to_sponge_bytes(UInt(len(x.encode()), 64)) + x.encode()         # :233-236
    if x is None:
        return b"\x00"                                                         # :317-322
    if isinstance(x, Some):
        return b"\x01" + to_sponge_bytes(x.item)
    if isinstance(x, WithLength):
        return to_sponge_bytes(UInt(len(x.item), 64)) + to_sponge_bytes(x.item)
    if isinstance(x, (list, tuple)):
        return b"".join(to_sponge_bytes(i) for i in x)                         # :41-48
    raise TypeError(type(x))


def to_sponge_field_elements(x, p: int) -> list[int]:
    x = _norm(x)
    if isinstance(x, bool):
        return [int(x)]
    if isinstance(x, UInt):
        return [x.value % p]
    if isinstance(x, SInt):
        return [x.value % p]                                                   # :195-201: -F::from(|v|) for v < 0
    if isinstance(x, Fe):
        if x.p != p:
            raise ValueError("Trying to absorb non-native field elements.")    # field_cast(..).unwrap(), :106-122
        return [x.value]
    if isinstance(x, TEPoint):
        if x.q != p:
            raise ValueError("Trying to absorb non-native field elements.")
        return [x.x, x.y]                                                      # :258-260
    if isinstance(x, (bytes, bytearray)):
        return bytes_to_field_elements(len(x).to_bytes(8, "little") + bytes(x), p)   # :137-141
    if isinstance(x, str):
        return to_sponge_field_elements(x.encode(), p)                         # :238-240
    if x is None:
        return [0]
    if isinstance(x, Some):
        return [1] + to_sponge_field_elements(x.item, p)
    if isinstance(x, WithLength):
        return to_sponge_field_elements(UInt(len(x.item), 64), p) + to_sponge_field_elements(x.item, p)
    if isinstance(x, (list, tuple)):
        out = []
        for i in x:
            out += to_sponge_field_elements(i, p)
        return out
    raise TypeError(type(x))


FULL = "Full"


def num_bits(size, p: int) -> int:
    """FieldElementSize::num_bits (R/sponge/mod.rs:38-48); size = FULL or an int (Truncated)."""
    if size == FULL:
        return p.bit_length() - 1
    if size > p.bit_length():
        raise ValueError("num_bits is greater than the capacity of the field.")
    return size


def squeeze_bytes(sponge, n: int) -> bytes:
    """R/sponge/poseidon/mod.rs:259-273."""
    p = sponge.cfg.p
    usable = (p.bit_length() - 1) // 8
    k = (n + usable - 1) // usable
    return b"".join(e.to_bytes(32, "little")[:usable] for e in sponge.squeeze_native_field_elements(k))[:n]


def squeeze_bits(sponge, n: int) -> list[bool]:
    """mod.rs:275-289."""
    p = sponge.cfg.p
    usable = p.bit_length() - 1
    k = (n + usable - 1) // usable
    bits = []
    for e in sponge.squeeze_native_field_elements(k):
        bits += [bool((e >> i) & 1) for i in range(usable)]
    return bits[:n]


def squeeze_field_elements_with_sizes(sponge, sizes, p2: int | None = None) -> list[int]:
    """mod.rs:291-307 over R/sponge/mod.rs:57-96, 170-187: native and all Full -> native squeeze; otherwise bits."""
    p = sponge.cfg.p
    p2 = p if p2 is None else p2
    if p2 == p and all(s == FULL for s in sizes):       # incl. empty `sizes`: permutes and enters Squeezing{0} (mod.rs:323-345)
        return sponge.squeeze_native_field_elements(len(sizes))
    if not sizes:                                       # non-native default implementation only (R/sponge/mod.rs:61-63)
        return []
    widths = [num_bits(s, p2) for s in sizes]
    bits = squeeze_bits(sponge, sum(widths))
    out, pos = [], 0
    for w in widths:
        v = sum(1 << i for i, b in enumerate(bits[pos:pos + w]) if b)
        pos += w
        out.append(v % p2)
    return out


def fork(sponge, domain: bytes):
    """CryptographicSponge::fork (R/sponge/mod.rs:145-153): clone, absorb len(domain) as usize bytes || domain."""
    import copy
    s = copy.deepcopy(sponge)
    s.absorb(to_sponge_field_elements(len(domain).to_bytes(8, "little") + bytes(domain), s.cfg.p))
    return s
