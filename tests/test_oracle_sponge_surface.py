"""Oracle-side restatement of the reference's sponge-interface tests (R/sponge/poseidon/tests.rs): squeeze order
independence (demo_bug :12-66), distinct encodings (:242-303), native squeeze cast (:305-319), the absorb!/collect
macros (:321-352), plus fork and the sized squeezes (R/sponge/mod.rs:57-96,145-153)."""
import random

import pytest

from helpers import oracle_config
from oracle import absorb as OA, poseidon as OP


def sponge(which="bls_default_r2"):
    _, cfg = oracle_config(which)
    return OP.PoseidonSponge(cfg), cfg


def absorb(s, x):
    s.absorb(OA.to_sponge_field_elements(x, s.cfg.p))


def test_demo_bug_squeeze_split_is_irrelevant():
    rnd = random.Random(1)
    _, cfg = oracle_config("bls_default_r2")
    inp = [rnd.randrange(cfg.p) for _ in range(3)]
    outs = []
    for split in ([3], [1, 1, 1], [2, 1], [1, 2]):
        s = OP.PoseidonSponge(cfg)
        s.absorb(inp)
        o = []
        for k in split:
            o += s.squeeze_native_field_elements(k)
        outs.append(o)
    assert outs[0] == outs[1] == outs[2] == outs[3]


def different(a, b):
    assert OA.to_sponge_bytes(a) != OA.to_sponge_bytes(b)
    s1, _ = sponge("bls_sponge_fixture")
    s2, _ = sponge("bls_sponge_fixture")
    absorb(s1, a)
    absorb(s2, b)
    assert s1.squeeze_native_field_elements(3) != s2.squeeze_native_field_elements(3)


def test_distinct_encodings():
    _, cfg = oracle_config("bls_sponge_fixture")
    rnd = random.Random(2)
    e = rnd.randrange(cfg.p - 1)
    different(OA.Fe(e, cfg.p), OA.Fe(e + 1, cfg.p))                                      # single_field_element
    l1 = [OA.Fe(rnd.randrange(cfg.p - 1), cfg.p) for _ in range(256)]
    l2 = list(l1)
    l2[3] = OA.Fe(l1[3].value + 1, cfg.p)
    different(l1, l2)                                                                    # list_with_constant_size_element
    different([OA.WithLength(bytes([1, 2, 3, 4])), OA.WithLength(bytes([5, 6]))],        # list_with_nonconstant_size_element
              [OA.WithLength(bytes([1, 2])), OA.WithLength(bytes([3, 4, 5, 6]))])


def test_squeeze_cast_native_and_sizes():
    s1, cfg = sponge("bls_sponge_fixture")
    absorb(s1, OA.Fe(12345, cfg.p))
    import copy
    s2, s3, s4 = copy.deepcopy(s1), copy.deepcopy(s1), copy.deepcopy(s1)
    native = s1.squeeze_native_field_elements(5)
    assert OA.squeeze_field_elements_with_sizes(s2, [OA.FULL] * 5) == native
    # truncated sizes go through the bit stream: element i = the next sizes[i] bits of squeeze_bits, little-endian
    sizes = [10, OA.FULL, 128, 1]
    got = OA.squeeze_field_elements_with_sizes(s3, sizes)
    bits = OA.squeeze_bits(s4, 10 + 254 + 128 + 1)
    pos = 0
    for g, w in zip(got, [10, 254, 128, 1]):
        assert g == sum(1 << i for i, b in enumerate(bits[pos:pos + w]) if b) % cfg.p
        pos += w
    with pytest.raises(ValueError):
        OA.squeeze_field_elements_with_sizes(s3, [256])
    # an empty native squeeze still permutes and enters Squeezing{0} (mod.rs:291-307 -> :323-345): the transcript moves on
    s5, _ = sponge("bls_sponge_fixture")
    absorb(s5, OA.Fe(7, cfg.p))
    s6 = copy.deepcopy(s5)
    assert OA.squeeze_field_elements_with_sizes(s5, []) == [] and s5.mode == ("S", 0)
    s6._permute()
    assert s5.state == s6.state
    # the non-native default implementation returns early without touching the sponge (R/sponge/mod.rs:61-63)
    s7 = copy.deepcopy(s6)
    from oracle import fields as OF
    assert OA.squeeze_field_elements_with_sizes(s7, [], OF.BN254_FR) == [] and s7.state == s6.state and s7.mode == s6.mode


def test_macros_shape():
    s1, cfg = sponge("bls_sponge_fixture")
    absorb(s1, [1, 2, 3, 4, 5, 6])
    absorb(s1, OA.Fe(114514, cfg.p))
    s2, _ = sponge("bls_sponge_fixture")
    for item in ([1, 2, 3, 4, 5, 6], OA.Fe(114514, cfg.p)):        # absorb!(s, a, b) absorbs each in turn (absorb.rs:346-352)
        absorb(s2, item)
    assert s1.squeeze_native_field_elements(3) == s2.squeeze_native_field_elements(3)
    assert OA.to_sponge_bytes([6, 5, 4, 3, 2, 1]) == b"".join(v.to_bytes(4, "little") for v in (6, 5, 4, 3, 2, 1))
    assert OA.to_sponge_field_elements([6, 5, -4], cfg.p) == [6, 5, cfg.p - 4]
    assert OA.to_sponge_bytes(OA.Fe(42, cfg.p)) == (42).to_bytes(32, "little")


def test_byte_strings_and_options():
    _, cfg = oracle_config("bls_default_r2")
    b = bytes(range(70))
    fe = OA.to_sponge_field_elements(b, cfg.p)
    raw = (70).to_bytes(8, "little") + b
    assert fe == [int.from_bytes(raw[i:i + 31], "little") for i in range(0, 78, 31)] and len(fe) == 3
    assert OA.to_sponge_field_elements("ab", cfg.p) == OA.to_sponge_field_elements(b"ab", cfg.p)
    assert OA.to_sponge_bytes("ab") == (2).to_bytes(8, "little") + b"ab"
    assert OA.to_sponge_field_elements(None, cfg.p) == [0]
    assert OA.to_sponge_field_elements(OA.Some(OA.UInt(9, 16)), cfg.p) == [1, 9]
    assert OA.to_sponge_bytes(OA.Some(OA.UInt(9, 16))) == b"\x01\x09\x00"
    with pytest.raises(ValueError):
        OA.to_sponge_field_elements(OA.Fe(1, cfg.p + 2), cfg.p)


def test_fork_separates_domains():
    s, cfg = sponge()
    s.absorb([1, 2, 3])
    a, b, a2 = OA.fork(s, b"alpha"), OA.fork(s, b"beta"), OA.fork(s, b"alpha")
    ra, rb = a.squeeze_native_field_elements(2), b.squeeze_native_field_elements(2)
    assert ra != rb and ra == a2.squeeze_native_field_elements(2)
    assert s.squeeze_native_field_elements(2) not in (ra, rb)            # the parent is untouched by fork
    assert len(OA.squeeze_bytes(OA.fork(s, b""), 100)) == 100
