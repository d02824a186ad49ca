"""Wire formats -- `CanonicalSerialize` / `CanonicalDeserialize` images of the types that cross the boundary:
PoseidonConfig (R/sponge/poseidon/mod.rs:25-45), pedersen::Parameters (R/crh/pedersen/mod.rs:28-31), merkle_tree::Path
and MultiPath (R/merkle_tree/mod.rs:139-152, 239-254), so a Rust process and this library can exchange parameters and
proofs as bytes.  R = /root/reference/crypto-primitives/src.

The derive macros write the struct fields in declaration order; the leaf encodings are ark-serialize / ark-ff / ark-ec
0.4 conventions (dependencies, absent from /root/reference), restated here from their published behaviour and NOT pinned
by any vector the reference holds:
  usize, u64        8 bytes little-endian (usize is written as u64)
  Vec<T>            u64 length, then the elements
  Fp (n bits)       ceil(n/8) bytes little-endian of the canonical (non-Montgomery) value
  TE affine point   compressed: y, with bit 7 of the last byte set when x > -x (x "negative");  uncompressed: x then y
  C: CurveGroup     written as its affine form
Host-side only; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

from .curves import TECurve
from .fields import Field


class Reader:
    def __init__(self, data: bytes):
        self.data, self.pos = bytes(data), 0

    def take(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise ValueError("unexpected end of input")            # SerializationError::IoError(UnexpectedEof)
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def u64(self) -> int:
        return int.from_bytes(self.take(8), "little")

    def done(self):
        if self.pos != len(self.data):
            raise ValueError("trailing bytes")


def u64(v: int) -> bytes:
    return int(v).to_bytes(8, "little")


def vec(items, enc) -> bytes:
    items = list(items)
    return u64(len(items)) + b"".join(enc(i) for i in items)


def read_vec(r: Reader, dec) -> list:
    return [dec(r) for _ in range(r.u64())]


# ---------------------------------------------------------------- field elements and points
def field_bytes(field: Field) -> int:
    return (field.modulus_bit_size + 7) // 8


def ser_field(field: Field, limbs) -> bytes:
    """(k, 4) or (4,) Montgomery limbs -> concatenated canonical encodings."""
    nb = field_bytes(field)
    return b"".join(v.to_bytes(nb, "little") for v in field.to_ints(limbs))


def de_field(field: Field, r: Reader) -> np.ndarray:
    v = int.from_bytes(r.take(field_bytes(field)), "little")
    if v >= field.modulus:
        raise ValueError("invalid data: field element not reduced")     # SerializationError::InvalidData
    return field.elements([v])[0]


def ser_point(curve: TECurve, xy, compress: bool = True) -> bytes:
    """(2, 4) affine Montgomery limbs."""
    f = curve.base_field
    x, y = f.to_ints(np.asarray(xy, dtype=np.uint64).reshape(2, 4))
    nb = field_bytes(f)
    if not compress:
        return x.to_bytes(nb, "little") + y.to_bytes(nb, "little")
    out = bytearray(y.to_bytes(nb, "little"))
    if x > (f.modulus - x) % f.modulus:
        out[-1] |= 0x80
    return bytes(out)


def de_point(curve: TECurve, r: Reader, compress: bool = True, validate: bool = True) -> np.ndarray:
    f = curve.base_field
    q, nb = f.modulus, field_bytes(f)
    if not compress:
        x, y = int.from_bytes(r.take(nb), "little"), int.from_bytes(r.take(nb), "little")
    else:
        raw = bytearray(r.take(nb))
        neg = bool(raw[-1] & 0x80)
        raw[-1] &= 0x7F
        y = int.from_bytes(raw, "little")
        if y >= q:
            raise ValueError("invalid data: coordinate not reduced")
        den = (-1 - curve.d * y * y) % q                           # a = -1:  x^2 = (1 - y^2) / (a - d y^2)
        x = curve._sqrt((1 - y * y) % q * pow(den, -1, q) % q) if den else None
        if x is None:
            raise ValueError("invalid data: not a curve point")
        if (x > (q - x) % q) != neg:
            x = (q - x) % q
    if x >= q or y >= q:
        raise ValueError("invalid data: coordinate not reduced")
    if validate:
        if not curve.is_on_curve((x, y)) or curve.mul(curve.scalar_modulus, (x, y)) != (0, 1):
            raise ValueError("invalid data: point not in the prime-order subgroup")
    return f.elements([x, y])


# ---------------------------------------------------------------- PoseidonConfig
def ser_poseidon_config(cfg) -> bytes:
    f = cfg.field
    t = cfg.rate + cfg.capacity
    ark = np.asarray(cfg.ark, dtype=np.uint64).reshape(-1, t, 4)
    mds = np.asarray(cfg.mds, dtype=np.uint64).reshape(t, t, 4)
    rows = lambda m: vec(m, lambda row: vec(row, lambda e: ser_field(f, e)))    # noqa: E731  Vec<Vec<F>>
    return u64(cfg.full_rounds) + u64(cfg.partial_rounds) + u64(cfg.alpha) + rows(ark) + rows(mds) + u64(cfg.rate) + u64(cfg.capacity)


def de_poseidon_config(field: Field, data: bytes):
    from .sponge.poseidon import PoseidonConfig
    r = Reader(data)
    rf, rp, alpha = r.u64(), r.u64(), r.u64()
    rows = lambda: read_vec(r, lambda r_: read_vec(r_, lambda r__: de_field(field, r__)))   # noqa: E731
    ark, mds = rows(), rows()
    rate, cap = r.u64(), r.u64()
    r.done()
    return PoseidonConfig(field, rf, rp, alpha, np.array(mds, dtype=np.uint64), np.array(ark, dtype=np.uint64), rate, cap)


# ---------------------------------------------------------------- pedersen::Parameters
def ser_pedersen_parameters(prm, compress: bool = True) -> bytes:
    """crh::pedersen::Parameters{generators: Vec<Vec<C>>} (the commitment's Parameters is not serialisable in the reference)."""
    return vec(prm.generators, lambda w: vec(w, lambda p: ser_point(prm.curve, p, compress)))


def de_pedersen_parameters(curve: TECurve, data: bytes, compress: bool = True, validate: bool = True):
    from .crh.pedersen import Parameters, Window
    r = Reader(data)
    gens = read_vec(r, lambda r_: read_vec(r_, lambda r__: de_point(curve, r__, compress, validate)))
    r.done()
    nw = len(gens)
    ws = len(gens[0]) if nw else 0
    if any(len(w) != ws for w in gens):
        raise ValueError("ragged generator table")
    return Parameters(curve, Window(ws, nw), np.array(gens, dtype=np.uint64).reshape(nw, ws, 2, 4))


# ---------------------------------------------------------------- Path / MultiPath
class FieldDigest:
    """Digest codec for trees whose digests are field elements (Poseidon)."""

    def __init__(self, field: Field):
        self.field = field

    def ser(self, d) -> bytes:
        return ser_field(self.field, d)

    def de(self, r: Reader):
        return de_field(self.field, r)


class PointDigest:
    """Digest codec for trees whose digests are affine points (Pedersen byte trees)."""

    def __init__(self, curve: TECurve, compress: bool = True, validate: bool = True):
        self.curve, self.compress, self.validate = curve, compress, validate

    def ser(self, d) -> bytes:
        return ser_point(self.curve, d, self.compress)

    def de(self, r: Reader):
        return de_point(self.curve, r, self.compress, self.validate)


def ser_path(path, leaf_codec, inner_codec=None) -> bytes:
    inner_codec = inner_codec or leaf_codec
    return leaf_codec.ser(path.leaf_sibling_hash) + vec(path.auth_path, inner_codec.ser) + u64(path.leaf_index)


def de_path(data: bytes, leaf_codec, inner_codec=None):
    from .merkle_tree import Path
    inner_codec = inner_codec or leaf_codec
    r = Reader(data)
    sib = leaf_codec.de(r)
    auth = read_vec(r, inner_codec.de)
    idx = r.u64()
    r.done()
    return Path(sib, auth, idx)


def ser_multipath(mp, leaf_codec, inner_codec=None) -> bytes:
    inner_codec = inner_codec or leaf_codec
    return (vec(mp.leaf_siblings_hashes, leaf_codec.ser) + vec(mp.auth_paths_prefix_lenghts, u64)
            + vec(mp.auth_paths_suffixes, lambda s: vec(s, inner_codec.ser)) + vec(mp.leaf_indexes, u64))


def de_multipath(data: bytes, leaf_codec, inner_codec=None):
    from .merkle_tree import MultiPath
    inner_codec = inner_codec or leaf_codec
    r = Reader(data)
    sibs = read_vec(r, leaf_codec.de)
    pre = read_vec(r, Reader.u64)
    suf = read_vec(r, lambda r_: read_vec(r_, inner_codec.de))
    idx = read_vec(r, Reader.u64)
    r.done()
    return MultiPath(sibs, pre, suf, idx)
