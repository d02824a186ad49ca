"""The device Poseidon code (csrc/poseidon.cuh: sparse-round schedule, lazy dot products, absorb
loop) compiled for the CPU with emulated PTX, driven by the product's own schedule derivation
(csrc/poseidon_host.hpp), against the oracle.  Bit-exact, both the sparse and the dense schedule."""
import ctypes as C
import random

import numpy as np
import pytest

from helpers import ALL_CONFIGS, build_host_shim, oracle_config, synth_elems
from oracle import cref, fields as OF, poseidon as OP

u64p = C.POINTER(C.c_uint64)
FID = {"bls12_381_fr": 0, "bn254_fr": 1, "jubjub_fr": 2, "bls12_377_fr": 3}


@pytest.fixture(scope="module", params=["merged", "split"])
def shim(request):
    """Both forms of the round loop (csrc/poseidon.cuh: one merged loop, or the partial rounds in a loop of their own --
    the library picks per field) for every field."""
    split = request.param == "split"
    return build_host_shim("poseidon_host_shim", defines=[f"CPB_POS_SPLIT={int(split)}"], tag="_" + request.param)


def _P(a):
    return a.ctypes.data_as(u64p)


def run(shim, fid, cfg, inp, allow_sparse):
    p = cfg.p
    ark = cref.ints_to_mont([x for r in cfg.ark for x in r], p)
    mds = cref.ints_to_mont([x for r in cfg.mds for x in r], p)
    n, L = inp.shape[0], inp.shape[1]
    out = np.zeros((n, 4), dtype=np.uint64)
    rc = shim.host_poseidon_crh(fid, cfg.rate, cfg.capacity, cfg.full_rounds, cfg.partial_rounds, C.c_ulonglong(cfg.alpha),
                                _P(ark), _P(mds), allow_sparse, _P(np.ascontiguousarray(inp)), C.c_long(L), C.c_long(n), _P(out))
    assert rc >= 0
    return rc, out


def check(shim, fid, cfg, Ls=(0, 1, 2, 3, 5), n=8, expect_sparse=None):
    O = cref.Poseidon(cfg)
    for L in Ls:
        inp = np.ascontiguousarray(synth_elems(100 + L, (n, max(L, 1)), cfg.p)[:, :L])
        if L:
            inp[0, :] = cref.ints_to_mont([cfg.p - 1] * L, cfg.p)
        exp = O.crh_batch(inp)
        for sp in (1, 0):
            rc, out = run(shim, fid, cfg, inp, sp)
            assert (out == exp).all(), (L, sp)
            if sp and expect_sparse is not None:
                assert rc == expect_sparse


@pytest.mark.parametrize("which", ALL_CONFIGS)
def test_named_configs(shim, which):
    fname, cfg = oracle_config(which)
    check(shim, FID[fname], cfg, expect_sparse=1)


@pytest.mark.parametrize("rate", [1, 3, 4, 8])
def test_other_widths(shim, rate):
    p = OF.BLS12_381_FR
    cfg = OP.get_default_poseidon_parameters(p, rate, False) if rate > 1 else None
    if cfg is None:
        rng = OF.SplitMix64(1)
        cfg = OP.PoseidonConfig(p, 4, 3, 17, [[rng.field(p) for _ in range(2)] for _ in range(7)],
                                [[rng.field(p) for _ in range(2)] for _ in range(2)], 1, 1)
    check(shim, 0, cfg, Ls=(0, 1, rate, rate + 1, 2 * rate + 1), n=4)


def test_adversarial_shapes(shim):
    """all-(p-1) matrices (largest lazy accumulations; singular -> dense fallback), capacity 2,
    no partial rounds, odd exponents."""
    rnd = random.Random(3)
    for p, fid in ((OF.BLS12_381_FR, 0), (OF.BN254_FR, 1)):
        for rate, cap, rf, rp, alpha, maxed in ((2, 1, 8, 5, 5, True), (2, 1, 2, 0, 3, False), (3, 2, 4, 2, 7, False),
                                                (2, 1, 8, 4, 31, False)):
            t = rate + cap
            ark = [[rnd.choice([p - 1, rnd.randrange(p)]) for _ in range(t)] for _ in range(rf + rp)]
            mds = [[p - 1] * t for _ in range(t)] if maxed else [[rnd.randrange(p) for _ in range(t)] for _ in range(t)]
            check(shim, fid, OP.PoseidonConfig(p, rf, rp, alpha, ark, mds, rate, cap), Ls=(0, 1, 2, 4), n=4,
                  expect_sparse=0 if (maxed or rp == 0) else 1)


def test_product_param_generation_matches_oracle(shim):
    """find_poseidon_ark_and_mds in the product's host code == oracle == reference KATs."""
    for fid, p, bits, rate, rf, rp in ((0, OF.BLS12_381_FR, 255, 2, 8, 31), (1, OF.BN254_FR, 254, 2, 8, 57),
                                       (0, OF.BLS12_381_FR, 255, 4, 8, 56)):
        t = rate + 1
        a = np.zeros(((rf + rp) * t, 4), dtype=np.uint64)
        m = np.zeros((t * t, 4), dtype=np.uint64)
        assert shim.host_poseidon_find_params(fid, rate, rf, rp, 0, _P(a), _P(m)) == 0
        ark, mds = OP.find_poseidon_ark_and_mds(p, bits, rate, rf, rp, 0)
        assert cref.mont_to_ints(a, p) == [x for r in ark for x in r]
        assert cref.mont_to_ints(m, p) == [x for r in mds for x in r]


def _cfg_arrays(cfg):
    p = cfg.p
    return (cref.ints_to_mont([x for r in cfg.ark for x in r], p), cref.ints_to_mont([x for r in cfg.mds for x in r], p))


@pytest.mark.parametrize("which,fid", [("bls_default_r2", 0), ("bn254_r2", 1)])
def test_sponge_absorb_then_squeeze_many(shim, which, fid):
    """absorb L, squeeze K native elements (R/sponge/poseidon/mod.rs:156-186,323-345) vs the oracle's state machine."""
    _, cfg = oracle_config(which)
    ark, mds = _cfg_arrays(cfg)
    for L, K in ((0, 1), (0, 5), (1, 2), (2, 3), (3, 3), (5, 7), (4, 1)):
        n = 5
        inp = np.ascontiguousarray(synth_elems(300 + L, (n, max(L, 1)), cfg.p)[:, :L])
        out = np.zeros((n, K, 4), dtype=np.uint64)
        assert shim.host_poseidon_sponge(fid, cfg.rate, cfg.capacity, cfg.full_rounds, cfg.partial_rounds, C.c_ulonglong(cfg.alpha),
                                         _P(ark), _P(mds), _P(inp), C.c_long(L), C.c_long(K), C.c_long(n), _P(out)) == 0
        ints = cref.mont_to_ints(inp, cfg.p)
        for i in range(n):
            s = OP.PoseidonSponge(cfg)
            s.absorb(ints[i * L:(i + 1) * L])
            assert cref.mont_to_ints(out[i], cfg.p) == s.squeeze_native_field_elements(K), (L, K)


def test_path_verification_device_code(shim):
    """Path::verify (R/merkle_tree/mod.rs:172-212) as run by one GPU thread, on the reference's field-tree shape."""
    from oracle import merkle as OM
    _, cfg = oracle_config("jubjub_merkle_fixture")
    ark, mds = _cfg_arrays(cfg)
    n, L = 16, 3
    lv = synth_elems(21, (n, L), cfg.p)
    li = cref.mont_to_ints(lv, cfg.p)
    h2 = lambda a, b: OP.two_to_one_compress(cfg, a, b)
    T = OM.MerkleTree.new([li[L * i:L * i + L] for i in range(n)], lambda l: OP.crh_evaluate(cfg, l), h2, h2)
    proofs = [T.generate_proof(i) for i in range(n)]
    plen = len(proofs[0][1])
    sib = cref.ints_to_mont([p[0] for p in proofs], cfg.p)
    paths = cref.ints_to_mont([x for p in proofs for x in p[1]], cfg.p)
    idx = np.arange(n, dtype=np.uint64)
    root = cref.ints_to_mont([T.root()], cfg.p)
    ok = np.zeros(n, dtype=np.uint8)

    def run(root_arr, idx_arr):
        assert shim.host_poseidon_verify(2, cfg.rate, cfg.capacity, cfg.full_rounds, cfg.partial_rounds, C.c_ulonglong(cfg.alpha), _P(ark),
                                         _P(mds), _P(root_arr), _P(np.ascontiguousarray(lv)), C.c_long(L), _P(sib), _P(paths), plen,
                                         idx_arr.ctypes.data_as(C.POINTER(C.c_ulonglong)), ok.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_long(n)) == 0
        return ok.copy()

    assert run(root, idx).all()
    assert not run(cref.ints_to_mont([(T.root() + 1) % cfg.p], cfg.p), idx).any()          # wrong root (tests/mod.rs:236-262)
    swapped = idx.copy(); swapped[[0, 1]] = swapped[[1, 0]]
    r = run(root, swapped)
    assert not r[0] and not r[1] and r[2:].all()                                            # wrong position


@pytest.mark.parametrize("which", ALL_CONFIGS)
def test_three_warp_team_schedule(shim, which):
    """CPU model of the latency-oriented team kernel (csrc/poseidon_team.cuh: one warp per state lane, lanes exchanged
    through a double-buffered array) == the oracle's two-to-one compression, sparse and dense schedules."""
    fname, cfg = oracle_config(which)
    p = cfg.p
    ark = cref.ints_to_mont([x for r in cfg.ark for x in r], p)
    mds = cref.ints_to_mont([x for r in cfg.mds for x in r], p)
    pairs = synth_elems(71, (40, 2), p)
    pairs[0] = cref.ints_to_mont([p - 1, p - 1], p)
    exp = cref.Poseidon(cfg).compress_batch(pairs)
    for sp in (1, 0):
        out = np.zeros((40, 4), dtype=np.uint64)
        rc = shim.host_poseidon_team_compress(FID[fname], cfg.full_rounds, cfg.partial_rounds, C.c_ulonglong(cfg.alpha), _P(ark), _P(mds), sp,
                                              _P(np.ascontiguousarray(pairs)), C.c_long(40), _P(out))
        assert rc == sp and np.array_equal(out, exp), (which, sp)


@pytest.mark.parametrize("which", ["bls_default_r2", "bn254_r2", "bls377_random"])
def test_crafted_sbox_operands(shim, which):
    """Inputs built so that the first S-box squares a value whose square has all-ones limbs (helpers.crafted_sbox_inputs):
    with the reduction bug fixed in round 1 (fp_sqr dropped a carry) about every fifth of these digests was wrong."""
    from helpers import crafted_sbox_inputs
    fname, cfg = oracle_config(which)
    inp = crafted_sbox_inputs(cfg, 600)
    exp = cref.Poseidon(cfg).crh_batch(inp)
    for sp in (1, 0):
        _, out = run(shim, FID[fname], cfg, inp, sp)
        assert (out == exp).all()


def test_odd_full_rounds(shim):
    """PoseidonConfig::new asserts shapes only (R/sponge/poseidon/mod.rs:189-217) and permute runs floor(RF/2) full rounds
    before and ceil(RF/2) after the partial ones (:98-121): odd RF must be accepted and match, sparse and dense, in the
    one-hash-per-thread code and in the three-warp team model."""
    rnd = random.Random(11)
    for p, fid in ((OF.BLS12_381_FR, 0), (OF.BN254_FR, 1)):
        for rf, rp, alpha in ((1, 0, 5), (1, 3, 5), (3, 4, 17), (5, 6, 5), (7, 2, 3), (3, 0, 5)):
            t = 3
            ark = [[rnd.randrange(p) for _ in range(t)] for _ in range(rf + rp)]
            mds = [[rnd.randrange(p) for _ in range(t)] for _ in range(t)]
            cfg = OP.PoseidonConfig(p, rf, rp, alpha, ark, mds, 2, 1)
            check(shim, fid, cfg, Ls=(0, 2, 3), n=4, expect_sparse=1 if (rp > 0 and rf >= 3) else 0)
            arkm = cref.ints_to_mont([x for r in ark for x in r], p)
            mdsm = cref.ints_to_mont([x for r in mds for x in r], p)
            pairs = synth_elems(5 + rf, (33, 2), p)
            exp = cref.Poseidon(cfg).compress_batch(pairs)
            for sp in (1, 0):
                out = np.zeros((33, 4), dtype=np.uint64)
                rc = shim.host_poseidon_team_compress(fid, rf, rp, C.c_ulonglong(alpha), _P(arkm), _P(mdsm), sp,
                                                      _P(np.ascontiguousarray(pairs)), C.c_long(33), _P(out))
                assert rc >= 0 and np.array_equal(out, exp), (rf, rp, sp)


def test_lazy_reduction_rounds_bn254_alpha5(shim):
    """BN254 Fr with alpha = 5 runs its rounds with unreduced lane values (poseidon.cuh LZ; bounds in tests/test_lazy_bounds.py):
    round constants at p-1 (the largest x = d + c), inputs at p-1 / 0 / 1, widths 2 and 3 (the widths LZ admits) and 4 (it does
    not), many random inputs -- sparse and dense schedules against the oracle."""
    rnd = random.Random(21)
    p = OF.BN254_FR
    for rate, rf, rp in ((2, 8, 57), (1, 8, 56), (3, 8, 56), (2, 2, 1), (2, 3, 2)):
        t = rate + 1
        for extreme in (True, False):
            ark = [[(p - 1 - rnd.randrange(3)) if extreme else rnd.randrange(p) for _ in range(t)] for _ in range(rf + rp)]
            mds = [[rnd.randrange(p) for _ in range(t)] for _ in range(t)]
            cfg = OP.PoseidonConfig(p, rf, rp, 5, ark, mds, rate, 1)
            n = 64 if rp > 10 else 256
            inp = synth_elems(300 + rate + rf, (n, rate), p)
            inp[0, :] = cref.ints_to_mont([p - 1] * rate, p)
            inp[1, :] = cref.ints_to_mont([0] * rate, p)
            inp[2, :] = cref.ints_to_mont([1] * rate, p)
            exp = cref.Poseidon(cfg).crh_batch(inp)
            for sp in (1, 0):
                rc, out = run(shim, 1, cfg, inp, sp)
                assert (out == exp).all(), (rate, rf, rp, extreme, sp)
