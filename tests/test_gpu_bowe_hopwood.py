"""GPU parity tests for the Bowe-Hopwood Pedersen CRH (R/crh/bowe_hopwood/mod.rs) through the C-ABI against the
oracle -- bit-exact.  PARITY UNPINNED w.r.t. the reference (it holds only a smoke test, mod.rs:253-271)."""
import numpy as np
import pytest

import crypto_primitives_b200 as cp
from crypto_primitives_b200.crh import bowe_hopwood as BH
from crypto_primitives_b200.crh.pedersen import Window
from oracle import bowe_hopwood as OBH, cref, fields as OF, jubjub as jj, pedersen as OPD

pytestmark = pytest.mark.gpu
_cache = {}


def setup(ws, nw, seed):
    key = (ws, nw, seed)
    if key not in _cache:
        ow = OPD.Window(ws, nw)
        oprm = OBH.setup(ow, seed)
        g = cp.BLS12_381_FR.elements([c for s in oprm.generators for pt in s for c in pt]).reshape(nw, ws, 2, 4)
        _cache[key] = (ow, oprm, cref.Pedersen(oprm, ow), BH.Parameters(cp.curves.JUBJUB, Window(ws, nw), g))
    return _cache[key]


@pytest.mark.parametrize("ws,nw,lens", [(63, 8, (189, 128, 64, 33, 2, 1, 0)),     # the reference's test window (mod.rs:258-262): 63 x 8 x 3 = 1512 bits
                                        (7, 3, (7, 4, 1)),                        # 63 bits: partial last group, tail chunks only
                                        (10, 4, (15, 8))])                        # group width does not divide the segment
def test_crh_matches_oracle(ws, nw, lens):
    ow, oprm, oc, prm = setup(ws, nw, 4)
    for ln in lens:
        n = 160
        inp = np.ascontiguousarray(cref.synth_bytes(300 + ln, n * max(ln, 1)).reshape(n, max(ln, 1))[:, :ln])
        if ln:
            inp[0, :] = 0xFF
            inp[1, :] = 0
        assert np.array_equal(BH.CRH.evaluate_batch(prm, inp), oc.bowe_hopwood_batch(inp, threads=8)), (ws, nw, ln)
    one = bytes(inp[5])
    assert cp.BLS12_381_FR.to_ints(BH.CRH.evaluate(prm, one))[0] == OBH.crh_evaluate(oprm, ow, one)       # python big-int oracle


def test_digest_depends_on_length_and_zero_chunks_count():
    """Only covered chunks contribute, and an all-zero chunk contributes its generator once (mod.rs:165-177)."""
    ow, oprm, oc, prm = setup(63, 8, 4)
    f = cp.BLS12_381_FR
    h1 = f.to_ints(BH.CRH.evaluate(prm, bytes(1)))[0]
    h2 = f.to_ints(BH.CRH.evaluate(prm, bytes(2)))[0]
    g = oprm.generators[0]
    assert h1 == jj.add(jj.add(g[0], g[1]), g[2])[0] and h1 != h2
    assert f.to_ints(BH.CRH.evaluate(prm, b""))[0] == 0                                                  # identity: x = 0


def test_two_to_one_and_length_guard():
    ow, oprm, oc, prm = setup(63, 8, 4)
    f = cp.BLS12_381_FR
    kids = cref.synth_field_mont(5, 2 * 64, jj.Q).reshape(64, 2, 4)
    got = BH.TwoToOneCRH.compress_batch(prm, kids)
    ints = cref.mont_to_ints(kids, jj.Q)
    for i in (0, 7, 63):
        assert f.to_ints(got[i])[0] == OBH.two_to_one_compress(oprm, ow, ints[2 * i], ints[2 * i + 1])
    l, r = bytes(range(20)), bytes(range(40, 60))
    assert f.to_ints(BH.TwoToOneCRH.evaluate(prm, l, r))[0] == OBH.two_to_one_evaluate(oprm, ow, l, r)
    with pytest.raises(ValueError):
        BH.CRH.evaluate_batch(prm, np.zeros((1, 190), dtype=np.uint8))                                   # 1520 bits > 1512
    with pytest.raises(ValueError):
        BH.CRH.setup(OF.SplitMix64(1), Window(64, 1))                                                    # > 62 chunks per segment for Jubjub


def test_setup_generators_and_throughput_shape():
    prm = BH.CRH.setup(OF.SplitMix64(9), Window(8, 2))
    f = cp.BLS12_381_FR
    g0 = tuple(f.to_ints(prm.generators[0, 0]))
    g1 = tuple(f.to_ints(prm.generators[0, 1]))
    assert cp.curves.JUBJUB.mul(16, g0) == g1                                                            # base, 16*base, ... (mod.rs:51-56)
    x = np.zeros((1 << 14, 6), dtype=np.uint8)
    x[:, 0] = np.arange(1 << 14) & 0xFF
    out = BH.CRH.evaluate_batch(prm, x)
    assert np.array_equal(out[:256], out[256:512]) and len({tuple(r) for r in out[:256]}) == 256


def test_merkle_tree_over_bowe_hopwood_digests():
    """MerkleTree<Config{LeafHash = TwoToOneHash = bowe_hopwood}> through the generic level loop of merkle_tree.Config:
    arrays against the oracle's tree over the big-integer Bowe-Hopwood restatement; proofs verify, a wrong leaf does not."""
    from crypto_primitives_b200.merkle_tree import BoweHopwoodByteConfig, MerkleTree
    from oracle import merkle as OM
    ow, oprm, oc, prm = setup(63, 9, 4)                  # 63 * 9 * 3 = 1701 bits >= 2 * 256-bit children + 8-byte leaves
    f = cp.BLS12_381_FR
    n = 32
    leaves = np.ascontiguousarray(cref.synth_bytes(77, n * 8).reshape(n, 8))
    cfg = BoweHopwoodByteConfig()
    tree = MerkleTree.new(prm, prm, leaves, cfg)
    comp = lambda l, r: OBH.two_to_one_compress(oprm, ow, l, r)          # noqa: E731
    otree = OM.MerkleTree.new([bytes(l) for l in leaves], lambda l: OBH.crh_evaluate(oprm, ow, l), comp, comp)
    assert f.to_ints(tree.leaf_nodes) == otree.leaf_nodes
    assert f.to_ints(tree.non_leaf_nodes) == otree.non_leaf_nodes
    root = tree.root()
    for i in (0, 13, 31):
        assert tree.generate_proof(i).verify(prm, prm, root, leaves[i], cfg)
    assert not tree.generate_proof(3).verify(prm, prm, root, leaves[4], cfg)
    assert tree.generate_multi_proof([1, 2, 30]).verify(prm, prm, root, leaves[[1, 2, 30]], cfg)
