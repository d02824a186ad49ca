"""Shared test helpers: oracle configs, fixture loading, host-shim builds."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import cref, fields as OF, poseidon as OP  # noqa: E402


def kats():
    return json.load(open(os.path.join(GOLDEN, "reference_kats.json")))


def fixture_config(name):
    j = json.load(open(os.path.join(GOLDEN, f"poseidon_fixture_{name}.json")))
    p = OF.MODULI[j["field"]]
    cfg = OP.PoseidonConfig(p, j["full_rounds"], j["partial_rounds"], j["alpha"],
                            [[int(x) % p for x in r] for r in j["ark"]],     # FromStr reduces mod p
                            [[int(x) % p for x in r] for r in j["mds"]], j["rate"], j["capacity"])
    return j["field"], cfg


_cfg_cache = {}


def oracle_config(which):
    """Named oracle PoseidonConfigs used across tests."""
    if which in _cfg_cache:
        return _cfg_cache[which]
    if which == "bls_default_r2":          # BASELINE configs 1, 2 (R/sponge/test.rs:15)
        v = ("bls12_381_fr", OP.get_default_poseidon_parameters(OF.BLS12_381_FR, 2, False))
    elif which == "bls_weights_r2":        # alpha = 257
        v = ("bls12_381_fr", OP.get_default_poseidon_parameters(OF.BLS12_381_FR, 2, True))
    elif which == "bn254_r2":              # BASELINE config 4
        ark, mds = OP.find_poseidon_ark_and_mds(OF.BN254_FR, 254, 2, 8, 57, 0)
        v = ("bn254_fr", OP.PoseidonConfig(OF.BN254_FR, 8, 57, 5, ark, mds, 2, 1))
    elif which == "jubjub_merkle_fixture":
        v = fixture_config("merkle_jubjub_fr")
    elif which == "bls_sponge_fixture":
        v = fixture_config("sponge_bls12_381_fr")
    elif which == "bls377_random":         # shape of R/crh/poseidon/constraints.rs:133-190 (alpha 31, random MDS/ARK)
        rng = OF.SplitMix64(377)
        p = OF.BLS12_377_FR
        v = ("bls12_377_fr", OP.PoseidonConfig(p, 8, 24, 31, [[rng.field(p) for _ in range(3)] for _ in range(32)],
                                               [[rng.field(p) for _ in range(3)] for _ in range(3)], 2, 1))
    else:
        raise KeyError(which)
    _cfg_cache[which] = v
    return v


ALL_CONFIGS = ["bls_default_r2", "bls_weights_r2", "bn254_r2", "jubjub_merkle_fixture", "bls_sponge_fixture", "bls377_random"]


def product_config(which):
    """The same parameters as a crypto_primitives_b200.PoseidonConfig."""
    import crypto_primitives_b200 as cp
    fname, cfg = oracle_config(which)
    return cp.PoseidonConfig.from_ints(cp.FIELDS[fname], cfg.full_rounds, cfg.partial_rounds, cfg.alpha, cfg.mds, cfg.ark,
                                       cfg.rate, cfg.capacity)


def synth_elems(seed, shape_elems, p):
    n = int(np.prod(shape_elems))
    return cref.synth_field_mont(seed, n, p).reshape(*shape_elems, 4)


def build_host_shim(name, defines=(), tag=""):
    """g++ build of tests/host/<name>.cpp (device code with PTX primitives emulated) -> CDLL."""
    src = os.path.join(ROOT, "tests", "host", name + ".cpp")
    out_dir = os.path.join(ROOT, "tests", "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, name + tag + ".so")
    deps = [src] + [os.path.join(ROOT, "crypto_primitives_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "crypto_primitives_b200", "csrc")) if f.endswith(("cuh", "hpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", *[f"-D{d}" for d in defines], "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def crafted_sbox_inputs(cfg, count, seed=1, limb64=False):
    """CRH inputs (count, 2, 4) in Montgomery limbs whose first-round S-box operand on lane `capacity` is a value whose
    512-bit square has all-ones limbs in its upper half -- the operands on which a reduction that drops a rare carry
    gives a wrong digest (tests/test_fp_host.py has the field-level versions).  The kernel squares the Montgomery
    representation s = in0 + ark[0][capacity] (both Montgomery), so in0 = s_crafted - ark."""
    import math
    import random
    rnd = random.Random(seed)
    p = cfg.p
    top_bits = (p * p).bit_length()
    ark_m = (cfg.ark[0][cfg.capacity] << 256) % p
    rows = []
    while len(rows) < count:
        run = rnd.choice((1, 1, 2))
        k = rnd.randrange(8, 15 - run + 1)
        if limb64:                                    # all-ones 64-bit limbs (the C oracle works on 4 x 64)
            run, k = 2, rnd.choice((8, 10, 12))
        if 32 * (k + run) >= top_bits - 2:
            continue
        v = rnd.randrange(1, (p * p) >> (32 * (k + run))) << (32 * (k + run))
        for j in range(run):
            v |= 0xFFFFFFFF << (32 * (k + j))
        v |= (0x80000000 | rnd.getrandbits(31)) << (32 * (k - 1))
        v |= rnd.getrandbits(32 * (k - 1))
        s = math.isqrt(v)
        if s >= p:
            continue
        in0 = (s - ark_m) % p                         # Montgomery limbs of the first input element
        rows.append([in0, rnd.randrange(p)])
    out = np.zeros((count, 2, 4), dtype=np.uint64)
    for i, (a, b) in enumerate(rows):
        for j, v in enumerate((a, b)):
            out[i, j] = [(v >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(4)]
    return out
