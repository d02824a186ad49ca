"""Pins the oracle (Python big-int and C restatement) to every known-answer constant the
reference's own tests hold for the Poseidon path (tests/golden/reference_kats.json, transcribed
from grain_lfsr.rs:190-218, traits.rs:163-358, sponge/poseidon/mod.rs:381-404)."""
import numpy as np
import pytest

from helpers import ALL_CONFIGS, kats, oracle_config, synth_elems
from oracle import cref, fields as OF, merkle as OM, poseidon as OP

K = kats()
P = OF.BLS12_381_FR


def test_modulus_matches_reference_test_field():
    assert int(K["field_modulus"]) == P                       # R/sponge/test.rs:6


def test_grain_lfsr_kat():
    g = K["grain_lfsr"]
    l = OP.PoseidonGrainLFSR(*g["args"])
    assert l.get_field_elements_rejection_sampling(1, P)[0] == int(g["rejection_sampling"][0])
    assert l.get_field_elements_rejection_sampling(1, P)[0] == int(g["rejection_sampling"][1])
    assert l.get_field_elements_mod_p(1, P)[0] == int(g["mod_p"][0])
    assert l.get_field_elements_mod_p(1, P)[0] == int(g["mod_p"][1])


@pytest.mark.parametrize("weights", [False, True])
@pytest.mark.parametrize("rate", range(2, 9))
def test_default_parameters_kat(rate, weights):
    e = K["default_params"]["weights" if weights else "constraints"][str(rate)]
    cfg = OP.get_default_poseidon_parameters(P, rate, weights)
    assert cfg.ark[0][0] == int(e["ark00"])
    assert cfg.mds[0][0] == int(e["mds00"])


def test_sponge_kat_python_and_c():
    s = K["sponge"]
    cfg = OP.get_default_poseidon_parameters(P, s["rate"], s["optimized_for_weights"])
    sp = OP.PoseidonSponge(cfg)
    sp.absorb([int(x) for x in s["absorb"]])
    assert sp.squeeze_native_field_elements(3) == [int(x) for x in s["squeeze3"]]
    # first squeezed element == CRH::evaluate([0,1,2]) (crh/poseidon/mod.rs:30-40), also on the C restatement
    assert OP.crh_evaluate(cfg, [0, 1, 2]) == int(s["squeeze3"][0])
    c = cref.Poseidon(cfg)
    out = c.crh_batch(cref.ints_to_mont([0, 1, 2], P).reshape(1, 3, 4))
    assert cref.mont_to_ints(out, P)[0] == int(s["squeeze3"][0])


def test_sponge_state_machine_vs_permutation_count():
    """absorb/squeeze bookkeeping (mod.rs:124-186): L elements at rate r cost max(1, ceil(L/r))
    permutations and compress(l, r) == CRH([l, r])."""
    _, cfg = oracle_config("bls_default_r2")
    assert OP.two_to_one_compress(cfg, 5, 7) == OP.crh_evaluate(cfg, [5, 7])
    st = OP.permute(cfg, [0, 5, 7])
    assert OP.crh_evaluate(cfg, [5, 7]) == st[1]
    st2 = OP.permute(cfg, [st[0], (st[1] + 9) % P, st[2]])
    assert OP.crh_evaluate(cfg, [5, 7, 9]) == st2[1]
    assert OP.crh_evaluate(cfg, []) == OP.permute(cfg, [0, 0, 0])[1]


@pytest.mark.parametrize("which", ALL_CONFIGS)
@pytest.mark.parametrize("L", [0, 1, 2, 3, 5])
def test_c_restatement_equals_python(which, L):
    _, cfg = oracle_config(which)
    n = 6
    x = np.ascontiguousarray(synth_elems(10 + L, (n, max(L, 1)), cfg.p)[:, :L])
    ints = cref.mont_to_ints(x, cfg.p)
    got = cref.mont_to_ints(cref.Poseidon(cfg).crh_batch(x, threads=2), cfg.p)
    assert got == [OP.crh_evaluate(cfg, ints[i * L:(i + 1) * L]) for i in range(n)]


@pytest.mark.parametrize("which", ["bls_default_r2", "bn254_r2"])
def test_c_restatement_on_crafted_carry_operands(which):
    """The C oracle against the big-integer oracle on inputs whose first S-box operand squares to all-ones limbs
    (32- and 64-bit aligned): the checker itself must not have a rare-carry defect."""
    from helpers import crafted_sbox_inputs
    _, cfg = oracle_config(which)
    for limb64 in (False, True):
        x = crafted_sbox_inputs(cfg, 40, seed=3, limb64=limb64)
        ints = cref.mont_to_ints(x, cfg.p)
        got = cref.mont_to_ints(cref.Poseidon(cfg).crh_batch(x, threads=2), cfg.p)
        assert got == [OP.crh_evaluate(cfg, ints[2 * i:2 * i + 2]) for i in range(40)]


def test_c_merkle_equals_python_heap_order():
    _, cfg = oracle_config("jubjub_merkle_fixture")
    n, L = 32, 3                                                  # R/merkle_tree/tests/mod.rs:293-296 shape
    lv = synth_elems(21, (n, L), cfg.p)
    c = cref.Poseidon(cfg)
    ln, nn = cref.poseidon_merkle(c, c, lv, threads=3)
    li = cref.mont_to_ints(lv, cfg.p)
    h2 = lambda a, b: OP.two_to_one_compress(cfg, a, b)
    T = OM.MerkleTree.new([li[L * i:L * i + L] for i in range(n)], lambda l: OP.crh_evaluate(cfg, l), h2, h2)
    assert cref.mont_to_ints(ln, cfg.p) == T.leaf_nodes
    assert cref.mont_to_ints(nn, cfg.p) == T.non_leaf_nodes
    # proofs verify, wrong root does not (R/merkle_tree/tests/mod.rs:226-262)
    for i in (0, 5, 31):
        pr = T.generate_proof(i)
        assert OM.verify_path(pr, T.leaf_nodes[i], T.root(), h2, h2)
        assert not OM.verify_path(pr, T.leaf_nodes[i], (T.root() + 1) % cfg.p, h2, h2)
    with pytest.raises(ValueError):
        cref.poseidon_merkle(c, c, lv[:24], threads=1)            # not a power of two


def test_multiproof_prefix_lengths_kat():
    _, cfg = oracle_config("bls_default_r2")
    h2 = lambda a, b: OP.two_to_one_compress(cfg, a, b)
    T = OM.MerkleTree(list(range(1, 9)), h2, h2)
    _, prefix, suffixes, idx = T.generate_multi_proof(range(8))
    assert prefix == K["multiproof_prefix_lengths_8_leaves"]["value"]     # R/merkle_tree/tests/mod.rs:166
    assert all(p + len(s) == 2 for p, s in zip(prefix, suffixes))


@pytest.mark.parametrize("fname", ["BLS12_381_FR", "BN254_FR", "JUBJUB_FR", "BLS12_377_FR"])
def test_c_field_routines_on_pattern_limbs(fname):
    """The C oracle's unrolled Montgomery product, dedicated squaring and branch-free add/sub against Python integers on
    operands whose 64-bit limbs are drawn from {0, 1, 2^64-1, 2^63, 2^32-1, random}: every carry position is hit."""
    p = getattr(OF, fname)
    rng = np.random.default_rng(7)
    pat = np.array([0, 1, (1 << 64) - 1, 1 << 63, (1 << 32) - 1, (1 << 64) - 2], dtype=np.uint64)
    n = 20000
    def draw():
        choice = rng.integers(0, 9, size=(n, 4))
        rnd = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        limbs = np.where(choice < 6, pat[np.minimum(choice, 5)], rnd)
        vals = [OF.from_limbs([int(x) for x in row]) % p for row in limbs]
        vals[0], vals[1], vals[2] = p - 1, 0, p - 2
        return vals
    A, B = draw(), draw()
    a = np.array([OF.to_limbs(v) for v in A], dtype=np.uint64)
    b = np.array([OF.to_limbs(v) for v in B], dtype=np.uint64)
    mul, sqr, add, sub = cref.field_ops(p, a, b)
    rinv = pow(1 << 256, -1, p)
    for i in range(n):
        assert OF.from_limbs([int(x) for x in mul[i]]) == A[i] * B[i] * rinv % p
        assert OF.from_limbs([int(x) for x in sqr[i]]) == A[i] * A[i] * rinv % p
        assert OF.from_limbs([int(x) for x in add[i]]) == (A[i] + B[i]) % p
        assert OF.from_limbs([int(x) for x in sub[i]]) == (A[i] - B[i]) % p
