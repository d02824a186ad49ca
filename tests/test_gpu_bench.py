"""GPU side of bench.py's parity chain: the on-device input generator equals the oracle's synthetic elements, and the
library reproduces the committed oracle results (tests/golden/bench_goldens.json) at the small sizes -- the same code
path bench.py runs at 2^20 / 2^22 / 2^24 and checks against the full-size entries of that file."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN, ROOT
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N
from oracle import cref, fields as OF

sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_inputs as BI  # noqa: E402

pytestmark = pytest.mark.gpu


def gold():
    return json.load(open(os.path.join(GOLDEN, "bench_goldens.json")))


@pytest.mark.parametrize("fname,seed", [("bn254_fr", BI.SEED_CONFIG4), ("bls12_381_fr", BI.SEED_CONFIG2), ("jubjub_fr", 77)])
def test_device_generator_equals_oracle_elements(fname, seed):
    import torch
    f = cp.FIELDS[fname]
    p = OF.MODULI[fname]
    exp = cref.synth_field_mont(seed, 5000, p)
    got = BI.field_elements_torch(torch, N, f.id, seed, 0, 5000, 0).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, exp)
    part = BI.field_elements_torch(torch, N, f.id, seed, 1234, 999, 0, chunk=256).cpu().numpy().view(np.uint64)
    assert np.array_equal(part, exp[1234:2233])


@pytest.mark.parametrize("key,fkey", [("small_merkle_2^10_poseidon_bn254", "bn254"), ("small_merkle_2^10_poseidon_bls12_381", "bls")])
def test_small_trees_reproduce_the_committed_oracle_results(key, fkey):
    import torch
    from crypto_primitives_b200.distributed import CudaPoseidonBackend
    g = gold()[key]
    prm = bench.poseidon_params(cp, fkey)
    n = 1 << g["log2_leaves"]
    leaves = BI.field_elements_torch(torch, N, prm.field.id, g["seed"], 0, 2 * n, 0).view(n, 2, 4)
    ln, nn = CudaPoseidonBackend(prm, prm, 0).build_local(leaves)
    torch.cuda.synchronize()
    assert bench.u64_list(nn[0].cpu()) == g["root"]
    assert [bench.u64_list(r) for r in nn[:31].cpu()] == g["top_nodes_heap_order"]
    for i, v in g["leaf_digest_samples"].items():
        assert bench.u64_list(ln[int(i)].cpu()) == v
    assert [int(x) for x in np.bitwise_xor.reduce(nn.cpu().numpy().view(np.uint64), axis=0)] == g["xor_of_all_nodes"]


def test_small_mixed_tree_and_sampled_pedersen_outputs():
    import torch
    from crypto_primitives_b200.distributed import CudaMixedBackend
    G = gold()
    g = G["small_mixed_merkle_2^8"]
    prm = bench.pedersen_setup(cp)
    node = bench.poseidon_params(cp, "bls")
    n = 1 << g["log2_leaves"]
    dev = torch.device("cuda", 0)
    leaves = BI.bytes_torch(torch, g["seed"], 0, 128 * n, dev).view(n, 128)
    ln, nn = CudaMixedBackend(prm, node, 0).build_local(leaves)
    torch.cuda.synchronize()
    assert bench.u64_list(nn[0].cpu()) == g["root"]
    for i, v in g["leaf_digest_samples"].items():
        assert bench.u64_list(ln[int(i)].cpu()) == v
    # config 3: the sampled inputs of the 2^20 batch, hashed / committed on their own
    gp = G["pedersen_2^20_jubjub"]
    idx = [int(i) for i in gp["crh_xy"]]
    inp = np.stack([BI.bytes_np(BI.SEED_CONFIG3, 128 * i, 128) for i in idx])
    rnd = np.stack([BI.randomness_np(BI.SEED_CONFIG3_RAND, i, 1)[0] for i in idx])
    from crypto_primitives_b200.commitment.pedersen import Commitment
    from crypto_primitives_b200.crh.pedersen import CRH
    h = CRH.evaluate_batch(prm, inp)
    c = Commitment.commit_batch(prm, inp, rnd)
    for k, i in enumerate(idx):
        assert [[int(x) for x in row] for row in h[k]] == gp["crh_xy"][str(i)]
        assert [[int(x) for x in row] for row in c[k]] == gp["commit_xy"][str(i)]


def test_launch_count_and_host_pinning_helpers():
    ctx = bench.poseidon_params(cp, "bn254").context(0)
    assert N.lib.cpb_merkle_poseidon_launch_count(ctx, 3) == 0
    assert 8 < N.lib.cpb_merkle_poseidon_launch_count(ctx, 1 << 24) < 200
    a = np.zeros(1 << 16, dtype=np.uint64)
    N.check(N.lib.cpb_host_register(a.ctypes.data, a.nbytes))
    N.check(N.lib.cpb_host_unregister(a.ctypes.data))
