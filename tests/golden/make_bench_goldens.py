"""Regenerates tests/golden/bench_goldens.json: ORACLE results (C restatement, oracle/cref) for the exact synthetic
workloads bench.py times, so that every driver-run bench line can state `root_matches_oracle` / `outputs_match_oracle`
without executing the oracle inside the GPU arm.  Inputs are the counter-based streams of bench_inputs.py,
identical at every GPU count.

    python tests/golden/make_bench_goldens.py [--threads T] [--only KEY ...]

CPU time on 8 cores: config 4 (2^24-leaf BN254 tree) ~4 min, config 5 (2^22 Pedersen leaves + Poseidon levels) ~3 min,
the rest seconds.  The `small_*` entries are the same pipelines at sizes the test-suite recomputes (tests/test_bench_goldens.py),
which pins this script to the committed JSON.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench_inputs as BI  # noqa: E402
from oracle import cref, fields as OF, pedersen as OPD, poseidon as OP  # noqa: E402

OUT = os.path.join(HERE, "bench_goldens.json")
SAMPLE_IDX = lambda n: sorted({0, 1, 2, 12345 % n, n // 2, n - 2, n - 1})  # noqa: E731


def poseidon_cfg(key):
    if key == "bls":
        return OP.get_default_poseidon_parameters(OF.BLS12_381_FR, 2, False)
    ark, mds = OP.find_poseidon_ark_and_mds(OF.BN254_FR, 254, 2, 8, 57, 0)
    return OP.PoseidonConfig(OF.BN254_FR, 8, 57, 5, ark, mds, 2, 1)


def field_leaves(seed, n_elems, p):
    """== bench_inputs.field_elements_torch(...) : raw stream limbs reduced mod p, Montgomery."""
    raw = BI.raw_field_limbs_np(seed, 0, n_elems)
    out = np.empty_like(raw)
    fld = cref.lib().oref_field_new(cref._p(cref.limbs(p)))
    cref.lib().oref_to_mont(fld, cref._p(out), cref._p(np.ascontiguousarray(raw)), n_elems)
    cref.lib().oref_field_free(fld)
    return out


def limbs_list(a):
    return [[int(x) for x in row] for row in np.asarray(a, dtype=np.uint64).reshape(-1, 4)]


def poseidon_tree(key, logn, seed, threads):
    cfg = poseidon_cfg(key)
    P = cref.Poseidon(cfg)
    n = 1 << logn
    leaves = field_leaves(seed, 2 * n, cfg.p).reshape(n, 2, 4)
    ln, nn = cref.poseidon_merkle(P, P, leaves, threads=threads)
    top = min(31, n - 1)
    return {"seed": seed, "log2_leaves": logn, "leaf_len": 2, "root": limbs_list(nn[0])[0],
            "top_nodes_heap_order": limbs_list(nn[:top]),                      # levels 0..4: subtree roots of 2/4/8/16-way splits (node 15 = root over the first 1/16 of the leaves)
            "leaf_digest_samples": {str(i): limbs_list(ln[i])[0] for i in SAMPLE_IDX(n)},
            "xor_of_all_nodes": [int(x) for x in np.bitwise_xor.reduce(nn, axis=0)]}


def permute_samples(logn, seed):
    cfg = poseidon_cfg("bls")
    P = cref.Poseidon(cfg)
    n = 1 << logn
    out = {}
    for i in SAMPLE_IDX(n):
        raw = BI.raw_field_limbs_np(seed, 3 * i, 3)
        st = np.empty_like(raw)
        fld = cref.lib().oref_field_new(cref._p(cref.limbs(cfg.p)))
        cref.lib().oref_to_mont(fld, cref._p(st), cref._p(np.ascontiguousarray(raw)), 3)
        cref.lib().oref_field_free(fld)
        out[str(i)] = limbs_list(P.permute(st))
    return {"seed": seed, "log2_states": logn, "t": 3, "state_samples": out}


def pedersen_params():
    w = OPD.Window(4, 256)
    return w, OPD.setup(w, BI.SEED_CONFIG3_PARAMS, commitment=True)


def pedersen_samples(logn):
    w, prm = pedersen_params()
    H = cref.Pedersen(prm, w)
    n = 1 << logn
    idx = SAMPLE_IDX(n)
    inp = np.stack([BI.bytes_np(BI.SEED_CONFIG3, 128 * i, 128) for i in idx])
    rnd = np.stack([BI.randomness_np(BI.SEED_CONFIG3_RAND, i, 1)[0] for i in idx])
    crh = H.batch(inp)
    com = H.batch(inp, rnd)
    g0 = prm.generators[0][0]
    return {"seed_inputs": BI.SEED_CONFIG3, "seed_randomness": BI.SEED_CONFIG3_RAND, "seed_params": BI.SEED_CONFIG3_PARAMS,
            "log2_inputs": logn, "input_len": 128, "window": [4, 256],
            "generator_0_0": [str(g0[0]), str(g0[1])], "randomness_generator_0": [str(c) for c in prm.randomness_generator[0]],
            "crh_xy": {str(i): limbs_list(crh[k]) for k, i in enumerate(idx)},
            "commit_xy": {str(i): limbs_list(com[k]) for k, i in enumerate(idx)}}


def mixed_tree(logn, threads):
    w, prm = pedersen_params()
    H = cref.Pedersen(OPD.Parameters(prm.generators), w)
    N = cref.Poseidon(poseidon_cfg("bls"))
    n = 1 << logn
    leaves = BI.bytes_np(BI.SEED_CONFIG5, 0, 128 * n).reshape(n, 128)
    ln, nn = cref.mixed_merkle(H, N, leaves, threads=threads)
    top = min(31, n - 1)
    return {"seed": BI.SEED_CONFIG5, "seed_params": BI.SEED_CONFIG3_PARAMS, "log2_leaves": logn, "leaf_len": 128,
            "root": limbs_list(nn[0])[0], "top_nodes_heap_order": limbs_list(nn[:top]),
            "leaf_digest_samples": {str(i): limbs_list(ln[i])[0] for i in SAMPLE_IDX(n)}}


JOBS = {
    "small_merkle_2^10_poseidon_bn254": lambda t: poseidon_tree("bn254", 10, BI.SEED_CONFIG4, t),
    "small_merkle_2^10_poseidon_bls12_381": lambda t: poseidon_tree("bls", 10, BI.SEED_CONFIG2, t),
    "small_mixed_merkle_2^8": lambda t: mixed_tree(8, t),
    "permute_2^22_bls12_381": lambda t: permute_samples(22, BI.SEED_CONFIG2_PERM),
    "pedersen_2^20_jubjub": lambda t: pedersen_samples(20),
    "merkle_2^20_poseidon_bls12_381": lambda t: poseidon_tree("bls", 20, BI.SEED_CONFIG2, t),
    "mixed_merkle_2^22": lambda t: mixed_tree(22, t),
    "merkle_2^24_poseidon_bn254": lambda t: poseidon_tree("bn254", 24, BI.SEED_CONFIG4, t),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--only", nargs="*")
    a = ap.parse_args()
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for key, fn in JOBS.items():
        if a.only and key not in a.only:
            continue
        t0 = time.time()
        gold[key] = fn(a.threads)
        gold[key]["oracle"] = "oracle/cref (C restatement), generated by tests/golden/make_bench_goldens.py"
        print(f"{key}: {time.time() - t0:.1f} s", flush=True)
        json.dump(gold, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
