"""Wire formats (crypto_primitives_b200/serialize.py): field order of the derived CanonicalSerialize impls
(R/sponge/poseidon/mod.rs:25-45, R/crh/pedersen/mod.rs:28-31, R/merkle_tree/mod.rs:139-152,239-254), round trips and
rejection of malformed input.  Host-side only.  The leaf encodings are ark-serialize conventions (unpinned, see module doc)."""
import numpy as np
import pytest

import crypto_primitives_b200 as cp
from crypto_primitives_b200 import serialize as S
from crypto_primitives_b200.crh.pedersen import Parameters, Window, create_generators, _points
from crypto_primitives_b200.merkle_tree import MultiPath, Path
from helpers import product_config
from oracle import fields as OF

F = cp.BLS12_381_FR
J = cp.curves.JUBJUB


def test_poseidon_config_layout_and_round_trip():
    cfg = product_config("bls_default_r2")            # t = 3, 8 + 31 rounds
    b = S.ser_poseidon_config(cfg)
    assert len(b) == 3 * 8 + (8 + 39 * (8 + 3 * 32)) + (8 + 3 * (8 + 3 * 32)) + 2 * 8
    assert b[:24] == (8).to_bytes(8, "little") + (31).to_bytes(8, "little") + (17).to_bytes(8, "little")
    assert b[24:32] == (39).to_bytes(8, "little") and b[32:40] == (3).to_bytes(8, "little")
    assert b[40:72] == F.to_ints(np.asarray(cfg.ark).reshape(-1, 4)[:1])[0].to_bytes(32, "little")     # canonical, not Montgomery
    assert b[-16:] == (2).to_bytes(8, "little") + (1).to_bytes(8, "little")
    c2 = S.de_poseidon_config(F, b)
    assert (c2.full_rounds, c2.partial_rounds, c2.alpha, c2.rate, c2.capacity) == (8, 31, 17, 2, 1)
    assert np.array_equal(np.asarray(c2.ark).reshape(-1), np.asarray(cfg.ark).reshape(-1))
    assert np.array_equal(np.asarray(c2.mds).reshape(-1), np.asarray(cfg.mds).reshape(-1))
    assert S.ser_poseidon_config(c2) == b
    with pytest.raises(ValueError):
        S.de_poseidon_config(F, b[:-1])
    with pytest.raises(ValueError):
        S.de_poseidon_config(F, b + b"\0")
    bad = bytearray(b)
    bad[40:72] = (F.modulus).to_bytes(32, "little")               # unreduced element
    with pytest.raises(ValueError):
        S.de_poseidon_config(F, bytes(bad))


@pytest.mark.parametrize("compress", [True, False])
def test_points_and_pedersen_parameters(compress):
    rng = OF.SplitMix64(11)
    w = Window(4, 3)
    gens = create_generators(J, w, rng)
    prm = Parameters(J, w, _points(J, [p for row in gens for p in row]).reshape(3, 4, 2, 4))
    b = S.ser_pedersen_parameters(prm, compress)
    assert len(b) == 8 + 3 * (8 + 4 * (32 if compress else 64))
    back = S.de_pedersen_parameters(J, b, compress)
    assert np.array_equal(back.generators, prm.generators) and (back.window.WINDOW_SIZE, back.window.NUM_WINDOWS) == (4, 3)
    # both signs of x survive compression; identity too
    x, y = gens[0][0]
    for pt in ((x, y), ((J.q - x) % J.q, y), (0, 1)):
        enc = S.ser_point(J, J.base_field.elements(list(pt)), compress)
        assert tuple(J.base_field.to_ints(S.de_point(J, S.Reader(enc), compress))) == pt
    if compress:
        assert (S.ser_point(J, J.base_field.elements([x, y]))[-1] ^ S.ser_point(J, J.base_field.elements([(J.q - x) % J.q, y]))[-1]) == 0x80
        with pytest.raises(ValueError):                            # a y with no x on the curve
            yy = 2
            while J._sqrt((1 - yy * yy) % J.q * pow((-1 - J.d * yy * yy) % J.q, -1, J.q) % J.q) is not None:
                yy += 1
            S.de_point(J, S.Reader(yy.to_bytes(32, "little")))
    else:
        with pytest.raises(ValueError):                            # off-curve pair
            S.de_point(J, S.Reader((1).to_bytes(32, "little") + (1).to_bytes(32, "little")), False)
    # a point of small order is on the curve but outside the prime-order subgroup
    with pytest.raises(ValueError):
        S.de_point(J, S.Reader(S.ser_point(J, J.base_field.elements([0, J.q - 1]), compress)), compress)


def test_path_and_multipath_round_trip():
    codec = S.FieldDigest(F)
    p = Path(F.elements([5])[0], [F.elements([i])[0] for i in (7, 8, 9)], 6)
    b = S.ser_path(p, codec)
    assert len(b) == 32 + 8 + 3 * 32 + 8 and b[-8:] == (6).to_bytes(8, "little")
    q = S.de_path(b, codec)
    assert q.leaf_index == 6 and F.to_ints(np.array(q.auth_path)) == [7, 8, 9] and F.to_ints(q.leaf_sibling_hash) == [5]
    mp = MultiPath([F.elements([1])[0], F.elements([2])[0]], [0, 2], [[F.elements([3])[0], F.elements([4])[0]], []], [4, 5])
    mb = S.ser_multipath(mp, codec)
    assert len(mb) == (8 + 64) + (8 + 16) + (8 + (8 + 64) + 8) + (8 + 16)
    m2 = S.de_multipath(mb, codec)
    assert m2.auth_paths_prefix_lenghts == [0, 2] and m2.leaf_indexes == [4, 5] and len(m2.auth_paths_suffixes[1]) == 0
    assert S.ser_multipath(m2, codec) == mb
    # point digests (Pedersen byte trees)
    pc = S.PointDigest(J)
    P = J.random_point(OF.SplitMix64(5))
    pp = Path(J.base_field.elements(list(P)), [J.base_field.elements(list(J.double(P)))], 1)
    assert np.array_equal(S.de_path(S.ser_path(pp, pc), pc).auth_path[0], pp.auth_path[0])


def test_c_encoders_equal_an_independent_python_restatement():
    """The library's encoders (csrc/cpb_serialize.cu through the C-ABI) against oracle/wire.py, byte for byte: PoseidonConfig, a
    real Merkle path / multiproof shape with field digests, points with both signs in both encodings."""
    from helpers import oracle_config
    from oracle import jubjub as jj, merkle as OM, wire as W
    _, ocfg = oracle_config("bls_default_r2")
    cfg = product_config("bls_default_r2")
    assert S.ser_poseidon_config(cfg) == W.poseidon_config(ocfg)
    p = ocfg.p
    rng = OF.SplitMix64(77)
    digests = [rng.field(p) for _ in range(16)]
    two = lambda l, r: (l * 3 + r * 5 + 1) % p                          # noqa: E731  any deterministic stand-in: layout test only
    t = OM.MerkleTree(digests, two, two)
    enc = lambda v: W.fe(v, p)                                          # noqa: E731
    codec = S.FieldDigest(F)
    for i in (0, 7, 15):
        sib, auth, idx = t.generate_proof(i)
        gp = Path(F.elements([sib])[0], [F.elements([a])[0] for a in auth], idx)
        assert S.ser_path(gp, codec) == W.path(sib, auth, idx, enc)
    sibs, prefix, suffixes, idx = t.generate_multi_proof([1, 2, 3, 9, 15])
    gm = MultiPath([F.elements([s])[0] for s in sibs], prefix, [[F.elements([a])[0] for a in suf] for suf in suffixes], idx)
    assert S.ser_multipath(gm, codec) == W.multipath(sibs, prefix, suffixes, idx, enc)
    back = S.de_multipath(S.ser_multipath(gm, codec), codec)
    assert back.auth_paths_prefix_lenghts == prefix and back.leaf_indexes == idx
    assert [[F.to_ints(a)[0] for a in suf] for suf in back.auth_paths_suffixes] == [list(s) for s in suffixes]
    for k in range(6):
        P = jj.mul(jj.COFACTOR, jj.point_from_y(rng.field(jj.Q)) or jj.IDENTITY)
        for pt in (P, ((-P[0]) % jj.Q, P[1])):
            for compress in (True, False):
                assert S.ser_point(J, J.base_field.elements(list(pt)), compress) == W.te_point(pt, jj.Q, compress)
