"""GPU parity tests for the field-leaf Merkle build (MerkleTree::new, R/merkle_tree/mod.rs:411-523)
and the proof / update surface the reference's own tests exercise (R/merkle_tree/tests/mod.rs:185-310)."""
import numpy as np
import pytest

from helpers import kats, oracle_config, product_config, synth_elems
import crypto_primitives_b200 as cp
from crypto_primitives_b200.merkle_tree import MerkleTree
from oracle import cref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,n,L", [("jubjub_merkle_fixture", 128, 3), ("bls_default_r2", 2, 2), ("bls_default_r2", 4, 1),
                                       ("bls_default_r2", 4096, 2), ("bn254_r2", 1 << 14, 2), ("bls_sponge_fixture", 64, 5)])
def test_build_matches_oracle_arrays(which, n, L):
    _, ocfg = oracle_config(which)
    cfg = product_config(which)
    leaves = synth_elems(300 + n, (n, L), ocfg.p)
    O = cref.Poseidon(ocfg)
    exp_leaf, exp_nodes = cref.poseidon_merkle(O, O, leaves, threads=8)
    t = MerkleTree.new(cfg, cfg, leaves)
    assert np.array_equal(t.leaf_nodes, exp_leaf)
    assert np.array_equal(t.non_leaf_nodes, exp_nodes)            # heap order, root at 0
    assert t.height() == n.bit_length() and np.array_equal(t.root(), exp_nodes[0])


def test_reference_field_mt_scenario():
    """field_mt_tests::good_root_test (R/merkle_tree/tests/mod.rs:289-309): 128 leaves x 3 elements over
    Jubjub Fr with the fixed parameters of test_utils.rs; proofs, multiproof, wrong root, updates."""
    _, ocfg = oracle_config("jubjub_merkle_fixture")
    cfg = product_config("jubjub_merkle_fixture")
    f = cfg.field
    leaves = synth_elems(9, (128, 3), ocfg.p)
    tree = MerkleTree.new(cfg, cfg, leaves)
    root = tree.root()
    for i in (0, 1, 2, 63, 64, 127):
        assert tree.generate_proof(i).verify(cfg, cfg, root, leaves[i])
    mp = tree.generate_multi_proof(range(128))
    assert mp.verify(cfg, cfg, root, leaves)
    wrong = f.elements([(f.to_ints(root)[0] + 1) % f.modulus])[0]              # root + F::one()
    assert not tree.generate_proof(0).verify(cfg, cfg, wrong, leaves[0])
    assert not mp.verify(cfg, cfg, wrong, leaves)
    upd = synth_elems(10, (5, 3), ocfg.p)
    for k, i in enumerate((2, 3, 5, 111, 127)):
        tree.update(i, upd[k])
        leaves[i] = upd[k]
    O = cref.Poseidon(ocfg)
    exp_leaf, exp_nodes = cref.poseidon_merkle(O, O, leaves, threads=8)
    assert np.array_equal(tree.non_leaf_nodes, exp_nodes) and np.array_equal(tree.leaf_nodes, exp_leaf)
    root = tree.root()
    for i in (0, 3, 111, 127):
        assert tree.generate_proof(i).verify(cfg, cfg, root, leaves[i])
    assert tree.check_update(7, upd[0], root) is False                          # tree untouched on failure
    assert np.array_equal(tree.non_leaf_nodes, exp_nodes)


def test_multiproof_prefix_lengths_kat():
    cfg = product_config("bls_default_r2")
    _, ocfg = oracle_config("bls_default_r2")
    tree = MerkleTree.new(cfg, cfg, synth_elems(4, (8, 2), ocfg.p))
    mp = tree.generate_multi_proof(range(8))
    assert mp.auth_paths_prefix_lenghts == kats()["multiproof_prefix_lengths_8_leaves"]["value"]
    # proofs survive the wire format (crypto_primitives_b200/serialize.py) and still verify against the GPU-built root
    from crypto_primitives_b200 import serialize as S
    codec = S.FieldDigest(cfg.field)
    leaves = synth_elems(4, (8, 2), ocfg.p)
    mp2 = S.de_multipath(S.ser_multipath(mp, codec), codec)
    assert mp2.verify(cfg, cfg, tree.root(), leaves)
    p5 = S.de_path(S.ser_path(tree.generate_proof(5), codec), codec)
    assert p5.verify(cfg, cfg, tree.root(), leaves[5]) and not p5.verify(cfg, cfg, tree.root(), leaves[4])
    cfg2 = S.de_poseidon_config(cfg.field, S.ser_poseidon_config(cfg))
    assert np.array_equal(MerkleTree.new(cfg2, cfg2, leaves).root(), tree.root())


def test_blank_and_new_with_leaf_digest():
    cfg = product_config("bls_default_r2")
    _, ocfg = oracle_config("bls_default_r2")
    O = cref.Poseidon(ocfg)
    t = MerkleTree.blank(cfg, cfg, 5)                                           # 16 zero digests
    z = np.zeros((16, 4), dtype=np.uint64)
    lvl = z
    while lvl.shape[0] > 1:
        lvl = O.compress_batch(lvl.reshape(-1, 2, 4))
    assert np.array_equal(t.root(), lvl[0]) and t.height() == 5
    d = synth_elems(6, (32,), ocfg.p)
    t2 = MerkleTree.new_with_leaf_digest(cfg, cfg, d)
    lvl = d
    while lvl.shape[0] > 1:
        lvl = O.compress_batch(lvl.reshape(-1, 2, 4))
    assert np.array_equal(t2.root(), lvl[0])


@pytest.mark.parametrize("n", [0, 1, 3, 24])
def test_not_power_of_two_is_rejected(n):
    """R/merkle_tree/mod.rs:430-433 asserts; the ABI returns CPB_NOT_POW2, the mirror raises."""
    cfg = product_config("bls_default_r2")
    with pytest.raises(ValueError):
        MerkleTree.new(cfg, cfg, np.zeros((n, 2, 4), dtype=np.uint64))


def test_full_size_tree_properties():
    """BASELINE config 2 (2^20 leaves, BLS12-381 Fr): the oracle recomputes the top levels and sampled
    bottom nodes; every level must be the compression of the level below (checked via the oracle on
    random positions), and the build must be deterministic."""
    _, ocfg = oracle_config("bls_default_r2")
    cfg = product_config("bls_default_r2")
    n = 1 << 20
    leaves = synth_elems(2, (n, 2), ocfg.p)
    t = MerkleTree.new(cfg, cfg, leaves)
    O = cref.Poseidon(ocfg)
    rng = np.random.default_rng(1)
    li = rng.choice(n, 256, replace=False)
    assert np.array_equal(t.leaf_nodes[li], O.crh_batch(leaves[li], threads=8))
    nodes = t.non_leaf_nodes
    idx = np.concatenate([np.arange(0, 1023), rng.choice(n // 2 - 1, 512, replace=False)])   # top 10 levels + samples
    kids = np.stack([nodes[2 * idx + 1], nodes[2 * idx + 2]], axis=1)
    assert np.array_equal(nodes[idx], O.compress_batch(kids, threads=8))
    bi = rng.choice(n // 2, 256, replace=False)                                           # bottom inner level
    kids = np.stack([t.leaf_nodes[2 * bi], t.leaf_nodes[2 * bi + 1]], axis=1)
    assert np.array_equal(nodes[n // 2 - 1 + bi], O.compress_batch(kids, threads=8))
    t2 = MerkleTree.new(cfg, cfg, leaves)
    assert np.array_equal(t2.non_leaf_nodes, nodes)


@pytest.mark.parametrize("which,logn", [("bls_default_r2", 20), ("bn254_r2", 22), ("bn254_r2", 24)])
def test_full_size_trees_equal_the_oracle_node_for_node(which, logn):
    """BASELINE configs 2 and 4 at full size: every leaf digest and every inner node against the C oracle run on all host
    threads (about 7 s / 30 s / 2 min of CPU).  A defect that hits one hash in 10^8 -- which sampled checks cannot see --
    changes the root here."""
    import os
    import time
    _, ocfg = oracle_config(which)
    cfg = product_config(which)
    n = 1 << logn
    O = cref.Poseidon(ocfg)
    threads = os.cpu_count() or 8
    probe = synth_elems(1, (1 << 15, 2), ocfg.p)                 # what the host can do: skip rather than run for an hour
    t0 = time.time()
    cref.poseidon_merkle(O, O, probe, threads=threads)
    predicted = (time.time() - t0) * (n / (1 << 15))
    if predicted > 400:
        pytest.skip(f"the C oracle would need ~{predicted:.0f} s on this host for 2^{logn} leaves")
    leaves = synth_elems(40 + logn, (n, 2), ocfg.p)
    t = MerkleTree.new(cfg, cfg, leaves)
    ln, nn = cref.poseidon_merkle(O, O, leaves, threads=threads)
    assert np.array_equal(t.non_leaf_nodes[0], nn[0]), "root differs"
    assert np.array_equal(t.leaf_nodes, ln)
    assert np.array_equal(t.non_leaf_nodes, nn)


def test_batched_path_verification():
    """All 4096 paths of a tree verified in one launch; tampered leaf / sibling / index / root are rejected."""
    _, ocfg = oracle_config("bls_default_r2")
    cfg = product_config("bls_default_r2")
    n = 4096
    leaves = synth_elems(12, (n, 2), ocfg.p)
    tree = MerkleTree.new(cfg, cfg, leaves)
    proofs = [tree.generate_proof(i) for i in range(n)]
    assert tree.verify_proofs_batch(proofs, leaves).all()
    bad_leaves = leaves.copy()
    bad_leaves[5, 0, 0] ^= np.uint64(1)
    r = tree.verify_proofs_batch(proofs, bad_leaves)
    assert not r[5] and r.sum() == n - 1
    proofs[9].leaf_index ^= 2
    proofs[11].auth_path[3] = proofs[12].auth_path[0]
    r = tree.verify_proofs_batch(proofs, leaves)
    assert not r[9] and not r[11] and r.sum() == n - 2
    wrong_root = cfg.field.elements([123])[0]
    assert not tree.verify_proofs_batch(proofs[:64], leaves[:64], wrong_root).any()
    # agrees with the single-path mirror of Path::verify
    assert proofs[100].verify(cfg, cfg, tree.root(), leaves[100])


def _oracle_tree(which, n, seed):
    _, ocfg = oracle_config(which)
    leaves = synth_elems(seed, (n, 2), ocfg.p)
    O = cref.Poseidon(ocfg)
    ln, nn = cref.poseidon_merkle(O, O, leaves, threads=8)
    return leaves, ln, nn


@pytest.mark.parametrize("which,world,logn", [("bn254_r2", 2, 6), ("bls_default_r2", 4, 12), ("bn254_r2", 8, 16), ("jubjub_merkle_fixture", 2, 18)])
def test_fused_root_exchange_with_all_ranks_on_one_gpu(which, world, logn):
    """The multi-GPU build's last kernel (k_poseidon_tree_top: local top levels + push of the root into every peer's
    exchange buffer + wait + replicated top) with all `world` ranks on THIS device, each on its own stream: the same code
    that runs over NVLink peer memory, testable on a one-GPU box.  Every rank's arrays against the oracle."""
    import ctypes as C
    import torch
    from crypto_primitives_b200 import _native as N
    from crypto_primitives_b200.distributed import level_slices
    cfg = product_config(which)
    n = 1 << logn
    per = n // world
    leaves, exp_leaf, exp_nodes = _oracle_tree(which, n, 4242 + logn)
    ctx = cfg.context(0)
    exs = (N.vp * world)()
    for r in range(world):
        h = N.vp()
        N.check(N.lib.cpb_exchange_create(0, world, r, C.byref(h)))
        exs[r] = h.value
    N.check(N.lib.cpb_exchange_connect_local(exs, world))
    try:
        dev = torch.device("cuda", 0)
        d_leaves = torch.from_numpy(np.ascontiguousarray(leaves).view(np.int64)).to(dev)
        streams = [torch.cuda.Stream(dev) for _ in range(world)]
        outs = []
        for rep in range(3):                                   # the epoch / parity logic: several collective calls in a row
            outs = []
            for r in range(world):
                ln = torch.zeros((per, 4), dtype=torch.int64, device=dev)
                nn = torch.zeros((per - 1, 4), dtype=torch.int64, device=dev)
                top = torch.zeros((world - 1, 4), dtype=torch.int64, device=dev)
                outs.append((ln, nn, top))
            torch.cuda.synchronize()
            for r in reversed(range(world)):                   # enqueue everything first, never wait in between
                N.check(N.lib.cpb_merkle_poseidon_build_sharded_dev(ctx, ctx, exs[r], d_leaves[r * per:(r + 1) * per].data_ptr(), 2, per,
                                                                    outs[r][0].data_ptr(), outs[r][1].data_ptr(), outs[r][2].data_ptr(),
                                                                    streams[r].cuda_stream))
            torch.cuda.synchronize()
        for r in range(world):
            ln, nn, top = (t.cpu().numpy().view(np.uint64) for t in outs[r])
            assert np.array_equal(ln, exp_leaf[r * per:(r + 1) * per])
            for gstart, cnt, lstart in level_slices(n, world, r):
                assert np.array_equal(nn[lstart:lstart + cnt], exp_nodes[gstart:gstart + cnt]), (r, gstart)
            assert np.array_equal(top, exp_nodes[:world - 1]), r
    finally:
        for r in range(world):
            N.lib.cpb_exchange_destroy(exs[r])


def test_tree_top_kernel_handles_every_small_shape():
    """from_digests for every n = 2 .. 2^14 power of two: the levels at or below 4096 hashes run in ONE launch of the
    four-warp tree-top kernel (per-CTA progress flags between levels); n > 8192 adds bulk launches above it."""
    cfg = product_config("bls_default_r2")
    _, ocfg = oracle_config("bls_default_r2")
    O = cref.Poseidon(ocfg)
    for logn in range(1, 15):
        n = 1 << logn
        d = synth_elems(900 + logn, (n,), ocfg.p)
        exp = []
        cur = d
        while len(cur) > 1:
            cur = O.compress_batch(cur.reshape(-1, 2, 4), threads=4)
            exp.insert(0, cur)
        exp = np.concatenate(exp)
        got = MerkleTree.new_with_leaf_digest(cfg, cfg, d).non_leaf_nodes
        assert np.array_equal(got, exp), logn


def test_level_synchronous_proof_and_update_batches_against_the_oracle():
    """R/merkle_tree/mod.rs:172-212 (Path::verify), :262-331 (MultiPath::verify with its look-up table), :627-726
    (update / check_update) as device batches -- height launches for any number of paths / leaves -- against the
    oracle's tree and proofs."""
    from crypto_primitives_b200.merkle_tree import MultiPath, PoseidonFieldConfig, verify_paths_batch
    from oracle import merkle as OM
    _, ocfg = oracle_config("jubjub_merkle_fixture")
    cfg = product_config("jubjub_merkle_fixture")
    f = cfg.field
    n = 256
    leaves = synth_elems(31, (n, 3), ocfg.p)
    tree = MerkleTree.new(cfg, cfg, leaves)
    O = cref.Poseidon(ocfg)
    exp_leaf, exp_nodes = cref.poseidon_merkle(O, O, leaves, threads=8)
    two = lambda l, r: tuple(int(x) for x in O.compress_batch(np.array([[l, r]], dtype=np.uint64))[0])   # noqa: E731
    otree = OM.MerkleTree([tuple(int(x) for x in d) for d in exp_leaf], two, two)
    root = tree.root()
    # --- paths generated by the ORACLE tree, verified by the device batch (generic level-synchronous form and the one-launch kernel)
    idxs = [0, 1, 2, 77, 128, 254, 255]
    proofs = []
    for i in idxs:
        sib, path, _ = otree.generate_proof(i)
        from crypto_primitives_b200.merkle_tree import Path
        proofs.append(Path(np.array(sib, dtype=np.uint64), [np.array(p, dtype=np.uint64) for p in path], i))
    ok = verify_paths_batch(cfg, cfg, root, leaves[idxs], proofs, PoseidonFieldConfig())
    assert ok.all() and tree.verify_proofs_batch(proofs, leaves[idxs]).all()
    bad = leaves[idxs].copy()
    bad[3, 1, 0] ^= np.uint64(1)
    proofs[5].auth_path[2] = proofs[5].auth_path[2].copy()
    proofs[5].auth_path[2][0] ^= np.uint64(4)
    ok = verify_paths_batch(cfg, cfg, root, bad, proofs, PoseidonFieldConfig())
    assert list(ok) == [True, True, True, False, True, False, True]
    # --- multiproof: the oracle's encoding == ours; verification hashes every tree node once per level
    for sel in (range(n), [3, 4, 5, 200], [255]):
        sibs, prefix, suffixes, idx = otree.generate_multi_proof(sel)
        mp = tree.generate_multi_proof(sel)
        assert mp.auth_paths_prefix_lenghts == prefix and mp.leaf_indexes == idx
        assert all(np.array_equal(a, np.array(b, dtype=np.uint64)) for a, b in zip(mp.leaf_siblings_hashes, sibs))
        assert mp.verify(cfg, cfg, root, leaves[idx])
        assert not mp.verify(cfg, cfg, f.elements([5])[0], leaves[idx])
        wrong = leaves[idx].copy()
        wrong[0, 0, 0] ^= np.uint64(1)
        assert not mp.verify(cfg, cfg, root, wrong)
    # Faithful to the reference's look-up table (mod.rs:304-306, `hash_lut.entry(..).or_insert_with(..)`): when both leaves of a
    # pair are in the multiproof, their parent is hashed from the FIRST one's (leaf, sibling hash) and the second leaf's
    # claimed hash is never used -- a tampered second leaf passes there, and therefore here.
    # The same table shadows whole subtrees: a later path stops being checked at its first ancestor that an earlier path
    # already inserted (from that path's own, correct, auth-path sibling).  Bug-compatible on purpose: same boolean as
    # the reference for every input.
    mp = tree.generate_multi_proof(range(n))
    second = leaves.copy()
    second[255, 0, 0] ^= np.uint64(1)
    assert mp.verify(cfg, cfg, root, second)
    shadowed = leaves.copy()
    shadowed[40, 0, 0] ^= np.uint64(1)                 # leaf 40: its ancestor at depth 2 was inserted by path 32
    assert mp.verify(cfg, cfg, root, shadowed)
    # --- k updates in one pass == k sequential updates == the oracle's tree of the final leaves
    upd_idx = [0, 1, 7, 100, 101, 255]
    new = synth_elems(32, (len(upd_idx), 3), ocfg.p)
    seq = MerkleTree.new(cfg, cfg, leaves)
    for k, i in enumerate(upd_idx):
        seq.update(i, new[k])
    tree.update_batch(upd_idx, new)
    final = leaves.copy()
    final[upd_idx] = new
    exp_leaf2, exp_nodes2 = cref.poseidon_merkle(O, O, final, threads=8)
    for t in (tree, seq):
        assert np.array_equal(t.leaf_nodes, exp_leaf2) and np.array_equal(t.non_leaf_nodes, exp_nodes2)
    # check_update: rejected -> untouched; accepted -> applied (single leaf and batch)
    again = synth_elems(33, (2, 3), ocfg.p)
    assert tree.check_update(9, again[0], root) is False and np.array_equal(tree.non_leaf_nodes, exp_nodes2)
    final[[9, 200]] = again
    _, exp_nodes3 = cref.poseidon_merkle(O, O, final, threads=8)
    assert tree.check_update_batch([9, 200], again, exp_nodes3[0]) is True
    assert np.array_equal(tree.non_leaf_nodes, exp_nodes3)


def test_vectorised_proof_arrays_equal_generate_proof():
    _, ocfg = oracle_config("bls_default_r2")
    cfg = product_config("bls_default_r2")
    for n in (2, 4, 512):
        leaves = synth_elems(44 + n, (n, 2), ocfg.p)
        tree = MerkleTree.new(cfg, cfg, leaves)
        idx = sorted({0, 1, n - 1, n // 2, (n * 3) // 7})
        sib, paths, ind = tree.generate_proofs_batch(idx)
        for k, i in enumerate(idx):
            p = tree.generate_proof(i)
            assert np.array_equal(sib[k], p.leaf_sibling_hash) and int(ind[k]) == i
            assert paths.shape[1] == len(p.auth_path) and all(np.array_equal(paths[k, j], p.auth_path[j]) for j in range(len(p.auth_path)))
        assert tree.verify_proofs_batch((sib, paths, ind), leaves[idx]).all()
        everything = tree.generate_proofs_batch(np.arange(n))
        assert tree.verify_proofs_batch(everything, leaves).all()
