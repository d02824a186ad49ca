"""Independent pure-Python encoders of the wire formats (test infrastructure only): the byte layouts of
csrc/cpb_serialize.cu restated over Python integers, so the library's C encoders / decoders are checked against a second
implementation and not only by round trips.  Same caveat as there: ark-serialize 0.4 conventions, unpinned by the reference."""
from __future__ import annotations


def u64(v: int) -> bytes:
    return int(v).to_bytes(8, "little")


def fe(v: int, p: int) -> bytes:
    return (v % p).to_bytes((p.bit_length() + 7) // 8, "little")


def vec(items, enc) -> bytes:
    items = list(items)
    return u64(len(items)) + b"".join(enc(i) for i in items)


def te_point(pt, q: int, compress: bool = True) -> bytes:
    x, y = pt
    if not compress:
        return fe(x, q) + fe(y, q)
    out = bytearray(fe(y, q))
    if x % q > (-x) % q:
        out[-1] |= 0x80
    return bytes(out)


def poseidon_config(cfg) -> bytes:
    """oracle.poseidon.PoseidonConfig -> bytes (field order of R/sponge/poseidon/mod.rs:26-45)."""
    rows = lambda m: vec(m, lambda row: vec(row, lambda e: fe(e, cfg.p)))     # noqa: E731
    return u64(cfg.full_rounds) + u64(cfg.partial_rounds) + u64(cfg.alpha) + rows(cfg.ark) + rows(cfg.mds) + u64(cfg.rate) + u64(cfg.capacity)


def path(leaf_sibling, auth_path, leaf_index, enc_leaf, enc_inner=None) -> bytes:
    """R/merkle_tree/mod.rs:146-152."""
    enc_inner = enc_inner or enc_leaf
    return enc_leaf(leaf_sibling) + vec(auth_path, enc_inner) + u64(leaf_index)


def multipath(sibs, prefix, suffixes, indexes, enc_leaf, enc_inner=None) -> bytes:
    """R/merkle_tree/mod.rs:245-254."""
    enc_inner = enc_inner or enc_leaf
    return vec(sibs, enc_leaf) + vec(prefix, u64) + vec(suffixes, lambda s: vec(s, enc_inner)) + vec(indexes, u64)
