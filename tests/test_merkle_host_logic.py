"""Host-side index logic of merkle_tree.py without a GPU: the level-synchronous batches (verify_paths_batch,
MultiPath.verify, update_batch / check_update, generate_proofs_batch, blank) driven by a toy Config whose hashes are
numpy arithmetic, against the oracle's sequential restatement of the reference (oracle/merkle.py) -- including the
reference's look-up-table semantics in MultiPath::verify under random tampering."""
import random

import numpy as np
import pytest

from crypto_primitives_b200.merkle_tree import Config, MerkleTree, MultiPath, Path, verify_paths_batch
from oracle import merkle as OM

M = np.uint64(0xFFFFFFFFFFFFFFFF)


class Toy(Config):
    """digest = 4 words; leaf = (L, 4) words."""
    digest_words = 4

    def leaf_hash_batch(self, prm, leaves, device):
        l = np.asarray(leaves, dtype=np.uint64)
        w = np.arange(1, l.shape[1] + 1, dtype=np.uint64).reshape(1, -1, 1)
        return (l * w * np.uint64(0x9E3779B97F4A7C15)).sum(axis=1) + np.uint64(17)

    def two_to_one_batch(self, prm, pairs, device):
        p = np.asarray(pairs, dtype=np.uint64).reshape(-1, 2, 4)
        x = p[:, 0] * np.uint64(0xBF58476D1CE4E5B9) + np.roll(p[:, 1], 1, axis=1) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(29))


CFG = Toy()
two = lambda l, r: tuple(int(v) for v in CFG.two_to_one_batch(None, np.array([[l, r]], dtype=np.uint64), 0)[0])   # noqa: E731


def build(n, seed):
    rng = np.random.default_rng(seed)
    leaves = rng.integers(0, 1 << 62, size=(n, 3, 4), dtype=np.uint64)
    tree = MerkleTree.new(None, None, leaves, config=CFG)
    digests = [tuple(int(v) for v in d) for d in CFG.leaf_hash_batch(None, leaves, 0)]
    return leaves, tree, OM.MerkleTree(digests, two, two)


def test_build_proofs_and_batches_equal_the_sequential_oracle():
    for n in (2, 4, 64):
        leaves, tree, otree = build(n, n)
        assert [tuple(int(v) for v in x) for x in tree.non_leaf_nodes] == otree.non_leaf_nodes
        root = tree.root()
        idx = sorted({0, 1, n - 1, n // 2})
        proofs = [tree.generate_proof(i) for i in idx]
        for p, i in zip(proofs, idx):
            sib, path, _ = otree.generate_proof(i)
            assert tuple(int(v) for v in p.leaf_sibling_hash) == sib and [tuple(int(v) for v in a) for a in p.auth_path] == path
            assert p.verify(None, None, root, leaves[i], config=CFG)
        assert verify_paths_batch(None, None, root, leaves[idx], proofs, CFG).all()
        sib, paths, ind = tree.generate_proofs_batch(idx)
        for k, p in enumerate(proofs):
            assert np.array_equal(sib[k], p.leaf_sibling_hash) and all(np.array_equal(paths[k, j], p.auth_path[j]) for j in range(len(p.auth_path)))
        bad = leaves[idx].copy()
        bad[0, 0, 0] ^= np.uint64(1)
        assert list(verify_paths_batch(None, None, root, bad, proofs, CFG)) == [False] + [True] * (len(idx) - 1)


def test_multipath_verify_is_the_reference_lut_algorithm_under_tampering():
    """Same boolean as the sequential algorithm of R/merkle_tree/mod.rs:262-331 for untouched and for randomly tampered
    inputs (leaves, sibling hashes, path suffixes, root) -- including the cases the look-up table lets through."""
    rnd = random.Random(5)
    n = 64
    leaves, tree, otree = build(n, 99)
    root = tree.root()
    accepted_tampered = 0
    for trial in range(300):
        k = rnd.choice([1, 2, 5, 17, n])
        sel = sorted(rnd.sample(range(n), k))
        mp = tree.generate_multi_proof(sel)
        lv = leaves[sel].copy()
        r = root.copy()
        what = rnd.choice(["none", "leaf", "sibling", "suffix", "root"])
        if what == "leaf":
            lv[rnd.randrange(k), rnd.randrange(3), rnd.randrange(4)] ^= np.uint64(1 << rnd.randrange(60))
        elif what == "sibling":
            j = rnd.randrange(k)
            mp.leaf_siblings_hashes[j] = mp.leaf_siblings_hashes[j].copy()
            mp.leaf_siblings_hashes[j][rnd.randrange(4)] ^= np.uint64(2)
        elif what == "suffix":
            cand = [j for j in range(k) if len(mp.auth_paths_suffixes[j])]
            if cand:
                j = rnd.choice(cand)
                q = rnd.randrange(len(mp.auth_paths_suffixes[j]))
                mp.auth_paths_suffixes[j][q] = mp.auth_paths_suffixes[j][q].copy()
                mp.auth_paths_suffixes[j][q][0] ^= np.uint64(8)
        elif what == "root":
            r[1] ^= np.uint64(1)
        digests = [tuple(int(v) for v in d) for d in CFG.leaf_hash_batch(None, lv, 0)]
        omp = ([tuple(int(v) for v in s) for s in mp.leaf_siblings_hashes], mp.auth_paths_prefix_lenghts,
               [[tuple(int(v) for v in a) for a in suf] for suf in mp.auth_paths_suffixes], mp.leaf_indexes)
        expect = OM.verify_multi_path(omp, digests, tuple(int(v) for v in r), two, two)
        got = mp.verify(None, None, r, lv, config=CFG)
        assert got == expect, (trial, what, sel)
        if what == "none":
            assert got
        elif got:
            accepted_tampered += 1
    assert accepted_tampered > 0        # the reference's table does let some tampered multiproofs through; so do we


def test_update_batch_equals_sequential_updates_and_rebuild():
    n = 128
    leaves, tree, _ = build(n, 7)
    rng = np.random.default_rng(8)
    idx = sorted(int(i) for i in rng.choice(n, 9, replace=False))
    new = rng.integers(0, 1 << 62, size=(len(idx), 3, 4), dtype=np.uint64)
    seq = MerkleTree.new(None, None, leaves, config=CFG)
    for k, i in enumerate(idx):
        seq.update(i, new[k])
    tree.update_batch(idx, new)
    final = leaves.copy()
    final[idx] = new
    fresh = MerkleTree.new(None, None, final, config=CFG)
    for t in (tree, seq):
        assert np.array_equal(t.leaf_nodes, fresh.leaf_nodes) and np.array_equal(t.non_leaf_nodes, fresh.non_leaf_nodes)
    before = tree.non_leaf_nodes.copy()
    assert tree.check_update(3, new[0], before[0]) is False and np.array_equal(tree.non_leaf_nodes, before)
    final[3] = new[0]
    fresh = MerkleTree.new(None, None, final, config=CFG)
    assert tree.check_update(3, new[0], fresh.root()) is True and np.array_equal(tree.non_leaf_nodes, fresh.non_leaf_nodes)
    with pytest.raises(AssertionError):
        tree.update_batch([1, 1], new[:2])


def test_blank_uses_the_default_digest():
    t = MerkleTree.blank(None, None, 4, config=CFG)
    assert t.leaf_nodes.shape == (8, 4) and not t.leaf_nodes.any()
    lvl = CFG.two_to_one_batch(None, np.zeros((4, 2, 4), dtype=np.uint64), 0)
    assert np.array_equal(t.non_leaf_nodes[3:7], lvl)


def test_multiproof_encoding_equals_the_oracle_and_the_reference_kat():
    """generate_multi_proof (vectorised) == the oracle's sequential restatement of mod.rs:589-623, and the prefix lengths of the
    all-leaves multiproof of an 8-leaf tree are the reference's own [0, 2, 1, 2, 0, 2, 1, 2] (R/merkle_tree/tests/mod.rs:166)."""
    leaves, tree, otree = build(8, 3)
    mp = tree.generate_multi_proof(range(8))
    assert mp.auth_paths_prefix_lenghts == [0, 2, 1, 2, 0, 2, 1, 2]
    rnd = random.Random(1)
    leaves, tree, otree = build(64, 4)
    for trial in range(40):
        sel = rnd.sample(range(64), rnd.choice([1, 2, 7, 33, 64])) + [5, 5]           # duplicates and disorder are normalised
        sibs, prefix, suffixes, idx = otree.generate_multi_proof(sel)
        mp = tree.generate_multi_proof(sel)
        assert mp.leaf_indexes == idx and mp.auth_paths_prefix_lenghts == prefix
        assert [tuple(int(v) for v in s) for s in mp.leaf_siblings_hashes] == sibs
        assert [[tuple(int(v) for v in a) for a in suf] for suf in mp.auth_paths_suffixes] == [list(s) for s in suffixes]
    assert tree.generate_multi_proof([]).leaf_indexes == []
    # smallest trees: a two-leaf tree has empty auth paths (height 2), a four-leaf tree one node per path
    for n in (2, 4):
        leaves, tree, otree = build(n, 10 + n)
        for sel in ([0], [1], list(range(n))):
            sibs, prefix, suffixes, idx = otree.generate_multi_proof(sel)
            mp = tree.generate_multi_proof(sel)
            assert mp.leaf_indexes == idx and mp.auth_paths_prefix_lenghts == prefix
            assert [tuple(int(v) for v in s) for s in mp.leaf_siblings_hashes] == sibs
            assert [[tuple(int(v) for v in a) for a in suf] for suf in mp.auth_paths_suffixes] == [list(s) for s in suffixes]
            assert mp.verify(None, None, tree.root(), leaves[idx], config=CFG)
