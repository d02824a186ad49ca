"""bench.py's synthetic inputs and committed oracle results (tests/golden/bench_goldens.json): the counter-based
generators of bench_inputs.py equal the oracle's sequential ones, the golden script reproduces the committed small
entries, and the two independent Pedersen set-ups (product mirror, oracle) derive the same generators."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN, ROOT
from oracle import cref, fields as OF, pedersen as OPD

sys.path.insert(0, ROOT)
import bench_inputs as BI  # noqa: E402


def gold():
    return json.load(open(os.path.join(GOLDEN, "bench_goldens.json")))


def test_counter_stream_equals_the_sequential_generator():
    seq = OF.SplitMix64(BI.SEED_CONFIG4)
    first = [seq.next() for _ in range(64)]
    assert [int(x) for x in BI.splitmix64_np(BI.SEED_CONFIG4, 0, 64)] == first
    assert [int(x) for x in BI.splitmix64_np(BI.SEED_CONFIG4, 17, 40)] == first[17:57]
    assert np.array_equal(BI.splitmix64_np(5, 0, 1000), cref.splitmix64_stream(5, 1000))
    assert np.array_equal(BI.bytes_np(9, 0, 1001), cref.synth_bytes(9, 1001))
    assert np.array_equal(BI.bytes_np(9, 131, 300), cref.synth_bytes(9, 431)[131:])


def test_torch_stream_equals_numpy_stream():
    import torch
    for seed, start, count in ((BI.SEED_CONFIG4, 0, 4096), (BI.SEED_CONFIG5, (1 << 33) + 5, 1000), (0xFFFFFFFFFFFFFFFF, 123, 77)):
        t = BI.splitmix64_torch(torch, seed, start, count, "cpu").numpy().view(np.uint64)
        assert np.array_equal(t, BI.splitmix64_np(seed, start, count))
    b = BI.bytes_torch(torch, BI.SEED_CONFIG3, 128 * 5, 128 * 3, "cpu").numpy()
    assert np.array_equal(b, BI.bytes_np(BI.SEED_CONFIG3, 128 * 5, 128 * 3))
    r = BI.randomness_torch(torch, BI.SEED_CONFIG3_RAND, 7, 9, "cpu").numpy()
    assert np.array_equal(r, BI.randomness_np(BI.SEED_CONFIG3_RAND, 7, 9))
    assert int(r[:, 31].max()) <= 7


def test_field_elements_are_the_oracle_synthetic_elements():
    p = OF.BN254_FR
    raw = BI.raw_field_limbs_np(BI.SEED_CONFIG4, 100, 50)
    vals = [OF.from_limbs([int(x) for x in row]) % p for row in raw]
    exp = cref.synth_field_mont(BI.SEED_CONFIG4, 150, p)[100:]
    assert cref.mont_to_ints(exp, p) == vals


def test_golden_script_reproduces_the_committed_small_entries():
    sys.path.insert(0, GOLDEN)
    import make_bench_goldens as M
    g = gold()
    for key in ("small_merkle_2^10_poseidon_bn254", "small_merkle_2^10_poseidon_bls12_381", "small_mixed_merkle_2^8"):
        fresh = M.JOBS[key](2)
        for k, v in fresh.items():
            assert g[key][k] == v, (key, k)
    for key in ("merkle_2^24_poseidon_bn254", "merkle_2^20_poseidon_bls12_381", "mixed_merkle_2^22", "permute_2^22_bls12_381",
                "pedersen_2^20_jubjub"):
        assert key in g, f"{key} missing: run tests/golden/make_bench_goldens.py"
    assert len(g["merkle_2^24_poseidon_bn254"]["top_nodes_heap_order"]) == 31


def test_product_and_oracle_pedersen_setups_agree():
    """Two independent implementations (crypto_primitives_b200.curves host arithmetic, oracle.jubjub) draw the same
    generators from the same stream -- bench.py uses the former, the golden script the latter."""
    import crypto_primitives_b200 as cp
    from crypto_primitives_b200.commitment.pedersen import Commitment
    from crypto_primitives_b200.crh.pedersen import Window
    prm = Commitment.setup(BI.StreamRng(BI.SEED_CONFIG3_PARAMS), Window(4, 256))
    o = OPD.setup(OPD.Window(4, 256), BI.SEED_CONFIG3_PARAMS, commitment=True)
    f = cp.BLS12_381_FR
    for w in (0, 1, 100, 255):
        for j in range(4):
            assert tuple(f.to_ints(prm.generators[w, j])) == tuple(o.generators[w][j])
    for k in (0, 1, 251):
        assert tuple(f.to_ints(prm.randomness_generator[k])) == tuple(o.randomness_generator[k])
    g = gold()["pedersen_2^20_jubjub"]
    assert [str(v) for v in f.to_ints(prm.generators[0, 0])] == g["generator_0_0"]


def test_bench_multiply_add_model_and_host_info():
    import bench
    assert bench.wide_madds_per_perm("bn254", 3, 8, 57, 5) == 61896
    assert bench.wide_madds_per_perm("bls", 3, 8, 31, 17) == 44784
    assert bench.wide_madds_per_perm("bn254", 3, 8, 57, 5, crh=True) == 61056      # minus one S-box and two last-round rows
    assert bench.wide_madds_per_perm("bls", 3, 8, 31, 17, crh=True) == 43856
    info = bench.host_cpu_info()
    assert 1 <= info["threads"] <= info["affinity"]
