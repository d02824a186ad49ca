"""Wire formats -- `CanonicalSerialize` / `CanonicalDeserialize` images of the types that cross the boundary:
PoseidonConfig (R/sponge/poseidon/mod.rs:25-45), pedersen::Parameters (R/crh/pedersen/mod.rs:28-31), merkle_tree::Path
and MultiPath (R/merkle_tree/mod.rs:139-152, 239-254), so a Rust process and this library can exchange parameters and
proofs as bytes.  R = /root/reference/crypto-primitives/src.

Thin wrappers: the encoders and decoders live in the library behind the C-ABI (include/cpb200.h "wire formats",
csrc/cpb_serialize.cu: host code, no GPU), so any host that binds the library has them.  The derive macros write the struct
fields in declaration order; the leaf encodings are ark-serialize / ark-ff / ark-ec 0.4 conventions (dependencies, absent
from /root/reference), restated from their published behaviour and NOT pinned by any vector the reference holds:
  usize, u64        8 bytes little-endian (usize is written as u64)
  Vec<T>            u64 length, then the elements
  Fp (n bits)       ceil(n/8) bytes little-endian of the canonical (non-Montgomery) value
  TE affine point   compressed: y, with bit 7 of the last byte set when x > -x (x "negative");  uncompressed: x then y
  C: CurveGroup     written as its affine form
Malformed input raises ValueError (the reference: SerializationError::{IoError(UnexpectedEof), InvalidData}).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .curves import TECurve
from .fields import Field


def _check(st: int):
    if st in (N.CPB_BAD_LENGTH, N.CPB_BAD_PARAMS):
        raise ValueError(N.lib.cpb_last_error().decode("utf-8", "replace"))
    N.check(st)


def _u8(b: bytes):
    return (C.c_uint8 * max(len(b), 1)).from_buffer_copy(bytes(b) + (b"\0" if not b else b""))


def _u64a(a):
    return np.ascontiguousarray(a, dtype=np.uint64).ctypes.data_as(N.u64p)


def _out(fn, *args) -> bytes:
    """Serialiser with (out, cap, *written) at the end: size query, then the bytes."""
    w = C.c_size_t()
    _check(fn(*args, None, 0, C.byref(w)))
    buf = (C.c_uint8 * max(w.value, 1))()
    _check(fn(*args, buf, w.value, C.byref(w)))
    return bytes(buf[:w.value])


class Reader:
    """Cursor over a byte string for callers that decode a sequence of leaf encodings themselves."""

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


# ---------------------------------------------------------------- field elements and points
def field_bytes(field: Field) -> int:
    return int(N.lib.cpb_field_serialized_size(field.id))


def ser_field(field: Field, limbs) -> bytes:
    """(k, 4) or (4,) Montgomery limbs -> concatenated canonical encodings."""
    a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    out = (C.c_uint8 * max(a.shape[0] * field_bytes(field), 1))()
    _check(N.lib.cpb_field_serialize(field.id, _u64a(a), a.shape[0], out))
    return bytes(out[:a.shape[0] * field_bytes(field)])


def de_field(field: Field, r: Reader) -> np.ndarray:
    raw = r.take(field_bytes(field))
    out = np.empty(4, dtype=np.uint64)
    _check(N.lib.cpb_field_deserialize(field.id, _u8(raw), 1, out.ctypes.data_as(N.u64p)))
    return out


def point_bytes(curve: TECurve, compress: bool = True) -> int:
    return int(N.lib.cpb_point_serialized_size(curve.id, int(compress)))


def ser_point(curve: TECurve, xy, compress: bool = True) -> bytes:
    """(2, 4) affine Montgomery limbs."""
    a = np.ascontiguousarray(xy, dtype=np.uint64).reshape(2, 4)
    out = (C.c_uint8 * point_bytes(curve, compress))()
    _check(N.lib.cpb_point_serialize(curve.id, _u64a(a), 1, int(compress), out))
    return bytes(out)


def de_point(curve: TECurve, r: Reader, compress: bool = True, validate: bool = True) -> np.ndarray:
    raw = r.take(point_bytes(curve, compress))
    out = np.empty((2, 4), dtype=np.uint64)
    _check(N.lib.cpb_point_deserialize(curve.id, _u8(raw), 1, int(compress), int(validate), out.ctypes.data_as(N.u64p)))
    return out


# ---------------------------------------------------------------- PoseidonConfig
def ser_poseidon_config(cfg) -> bytes:
    t = cfg.rate + cfg.capacity
    ark = np.ascontiguousarray(cfg.ark, dtype=np.uint64).reshape(-1, t, 4)
    mds = np.ascontiguousarray(cfg.mds, dtype=np.uint64).reshape(t, t, 4)
    return _out(N.lib.cpb_poseidon_config_serialize, cfg.field.id, cfg.rate, cfg.capacity, cfg.full_rounds, cfg.partial_rounds,
                cfg.alpha, _u64a(ark), _u64a(mds))


def de_poseidon_config(field: Field, data: bytes):
    from .sponge.poseidon import PoseidonConfig
    rate, cap, rf, rp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    alpha = C.c_uint64()
    buf = _u8(data)
    _check(N.lib.cpb_poseidon_config_deserialize(field.id, buf, len(data), C.byref(rate), C.byref(cap), C.byref(rf), C.byref(rp),
                                                 C.byref(alpha), None, 0, None, 0))
    t = rate.value + cap.value
    ark = np.empty((rf.value + rp.value, t, 4), dtype=np.uint64)
    mds = np.empty((t, t, 4), dtype=np.uint64)
    _check(N.lib.cpb_poseidon_config_deserialize(field.id, buf, len(data), None, None, None, None, None, ark.ctypes.data_as(N.u64p),
                                                 ark.size // 4, mds.ctypes.data_as(N.u64p), mds.size // 4))
    return PoseidonConfig(field, rf.value, rp.value, alpha.value, mds, ark, rate.value, cap.value)


# ---------------------------------------------------------------- pedersen::Parameters
def ser_pedersen_parameters(prm, compress: bool = True) -> bytes:
    """crh::pedersen::Parameters{generators: Vec<Vec<C>>} (the commitment's Parameters is not serialisable in the reference)."""
    g = np.ascontiguousarray(prm.generators, dtype=np.uint64)
    return _out(N.lib.cpb_pedersen_parameters_serialize, prm.curve.id, prm.window.WINDOW_SIZE, prm.window.NUM_WINDOWS, _u64a(g), int(compress))


def de_pedersen_parameters(curve: TECurve, data: bytes, compress: bool = True, validate: bool = True):
    from .crh.pedersen import Parameters, Window
    ws, nw = C.c_int(), C.c_int()
    buf = _u8(data)
    _check(N.lib.cpb_pedersen_parameters_deserialize(curve.id, buf, len(data), int(compress), int(validate), C.byref(ws), C.byref(nw), None, 0))
    g = np.empty((nw.value, ws.value, 2, 4), dtype=np.uint64)
    _check(N.lib.cpb_pedersen_parameters_deserialize(curve.id, buf, len(data), int(compress), 0, None, None, g.ctypes.data_as(N.u64p),
                                                     nw.value * ws.value))
    return Parameters(curve, Window(ws.value, nw.value), g)


# ---------------------------------------------------------------- Path / MultiPath
class FieldDigest:
    """Digest codec for trees whose digests are field elements (Poseidon)."""

    def __init__(self, field: Field):
        self.field = field
        self.kind, self.id, self.words, self.validate = 0, field.id, 4, True

    def ser(self, d) -> bytes:
        return ser_field(self.field, d)

    def de(self, r: Reader):
        return de_field(self.field, r)


class PointDigest:
    """Digest codec for trees whose digests are affine points (Pedersen byte trees)."""

    def __init__(self, curve: TECurve, compress: bool = True, validate: bool = True):
        self.curve, self.compress, self.validate = curve, compress, validate
        self.kind, self.id, self.words = (1 if compress else 2), curve.id, 8

    def ser(self, d) -> bytes:
        return ser_point(self.curve, d, self.compress)

    def de(self, r: Reader):
        return de_point(self.curve, r, self.compress, self.validate)


def ser_path(path, leaf_codec, inner_codec=None) -> bytes:
    inner_codec = inner_codec or leaf_codec
    auth = np.ascontiguousarray(np.stack(path.auth_path), dtype=np.uint64) if len(path.auth_path) else np.zeros(inner_codec.words, dtype=np.uint64)
    return _out(N.lib.cpb_path_serialize, leaf_codec.kind, leaf_codec.id, inner_codec.kind, inner_codec.id, _u64a(path.leaf_sibling_hash),
                _u64a(auth), len(path.auth_path), int(path.leaf_index))


def de_path(data: bytes, leaf_codec, inner_codec=None):
    from .merkle_tree import Path
    inner_codec = inner_codec or leaf_codec
    sib = np.empty(leaf_codec.words, dtype=np.uint64)
    auth = np.empty((64, inner_codec.words), dtype=np.uint64)
    n, idx = C.c_size_t(), C.c_uint64()
    _check(N.lib.cpb_path_deserialize(leaf_codec.kind, leaf_codec.id, inner_codec.kind, inner_codec.id, int(leaf_codec.validate and inner_codec.validate),
                                      _u8(data), len(data), sib.ctypes.data_as(N.u64p), auth.ctypes.data_as(N.u64p), 64, C.byref(n),
                                      C.byref(idx)))
    shape = (lambda a, w: a.reshape(2, 4) if w == 8 else a)
    return Path(shape(sib, leaf_codec.words), [shape(auth[i].copy(), inner_codec.words) for i in range(n.value)], int(idx.value))


def ser_multipath(mp, leaf_codec, inner_codec=None) -> bytes:
    inner_codec = inner_codec or leaf_codec
    n = len(mp.leaf_indexes)
    sib = np.ascontiguousarray(np.stack(mp.leaf_siblings_hashes), dtype=np.uint64) if n else np.zeros(leaf_codec.words, dtype=np.uint64)
    flat = [d for s in mp.auth_paths_suffixes for d in s]
    suf = np.ascontiguousarray(np.stack(flat), dtype=np.uint64) if flat else np.zeros(inner_codec.words, dtype=np.uint64)
    pre = np.array(mp.auth_paths_prefix_lenghts, dtype=np.uint64)
    lens = np.array([len(s) for s in mp.auth_paths_suffixes], dtype=np.uint64)
    idx = np.array(mp.leaf_indexes, dtype=np.uint64)
    return _out(N.lib.cpb_multipath_serialize, leaf_codec.kind, leaf_codec.id, inner_codec.kind, inner_codec.id, n, _u64a(sib), _u64a(pre),
                _u64a(lens), _u64a(suf), _u64a(idx))


def de_multipath(data: bytes, leaf_codec, inner_codec=None):
    from .merkle_tree import MultiPath
    inner_codec = inner_codec or leaf_codec
    buf = _u8(data)
    n, ns = C.c_size_t(), C.c_size_t()
    v = int(leaf_codec.validate and inner_codec.validate)
    _check(N.lib.cpb_multipath_deserialize(leaf_codec.kind, leaf_codec.id, inner_codec.kind, inner_codec.id, v, buf, len(data), C.byref(n),
                                           C.byref(ns), None, None, None, None, None, 0, 0))
    sib = np.empty((max(n.value, 1), leaf_codec.words), dtype=np.uint64)
    suf = np.empty((max(ns.value, 1), inner_codec.words), dtype=np.uint64)
    pre, lens, idx = (np.empty(max(n.value, 1), dtype=np.uint64) for _ in range(3))
    _check(N.lib.cpb_multipath_deserialize(leaf_codec.kind, leaf_codec.id, inner_codec.kind, inner_codec.id, 0, buf, len(data), C.byref(n),
                                           C.byref(ns), sib.ctypes.data_as(N.u64p), pre.ctypes.data_as(N.u64p), lens.ctypes.data_as(N.u64p),
                                           suf.ctypes.data_as(N.u64p), idx.ctypes.data_as(N.u64p), n.value, ns.value))
    shape = (lambda a, w: a.reshape(2, 4) if w == 8 else a)
    suffixes, at = [], 0
    for i in range(n.value):
        k = int(lens[i])
        suffixes.append([shape(suf[at + j].copy(), inner_codec.words) for j in range(k)])
        at += k
    return MultiPath([shape(sib[i].copy(), leaf_codec.words) for i in range(n.value)], [int(x) for x in pre[:n.value]], suffixes,
                     [int(x) for x in idx[:n.value]])
