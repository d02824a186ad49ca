"""Regenerates tests/golden/*.json from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box, so the JSON files are committed).

  reference_kats.json       -- every known-answer constant the reference's own tests hold for the
                               Poseidon path (transcribed, with file:line).
  poseidon_fixture_*.json   -- the fixed parameter sets the reference's tests hash with
                               (decimal strings exactly as written; reduced mod p by the loader,
                               as ark-ff's FromStr does).
"""
import json
import os
import re

REF = "/root/reference/crypto-primitives/src"
OUT = os.path.dirname(os.path.abspath(__file__))


def decimals(path, lo=None, hi=None):
    src = open(path).read().split("\n")
    if lo is not None:
        src = src[lo - 1:hi]
    return re.findall(r'"(\d+)"', "\n".join(src))


def main():
    # --- merkle_tree/tests/test_utils.rs: Jubjub-Fr fixture, RF=8 RP=29 alpha=17, 0/1 near-MDS (:7-9, :647-651)
    d = decimals(f"{REF}/merkle_tree/tests/test_utils.rs")
    assert len(d) == 37 * 3, len(d)
    json.dump({"source": "R/merkle_tree/tests/test_utils.rs:6-653", "field": "jubjub_fr", "full_rounds": 8,
               "partial_rounds": 29, "alpha": 17, "rate": 2, "capacity": 1,
               "ark": [d[3 * i:3 * i + 3] for i in range(37)],
               "mds": [["1", "0", "1"], ["1", "1", "0"], ["0", "1", "1"]]},
              open(f"{OUT}/poseidon_fixture_merkle_jubjub_fr.json", "w"), indent=0)
    # --- sponge/poseidon/tests.rs:354-1054: BLS12-381 Fr fixture, alpha=17, 8 full + 29 partial
    d = decimals(f"{REF}/sponge/poseidon/tests.rs", 354, 1054)
    assert len(d) == 9 + 37 * 3, len(d)
    mds, ark = d[:9], d[9:]
    json.dump({"source": "R/sponge/poseidon/tests.rs:354-1054", "field": "bls12_381_fr", "full_rounds": 8,
               "partial_rounds": 29, "alpha": 17, "rate": 2, "capacity": 1,
               "ark": [ark[3 * i:3 * i + 3] for i in range(37)],
               "mds": [mds[3 * i:3 * i + 3] for i in range(3)]},
              open(f"{OUT}/poseidon_fixture_sponge_bls12_381_fr.json", "w"), indent=0)
    # --- KATs
    g = decimals(f"{REF}/sponge/poseidon/grain_lfsr.rs", 190, 218)
    s = decimals(f"{REF}/sponge/poseidon/mod.rs", 381, 404)
    t = decimals(f"{REF}/sponge/poseidon/traits.rs", 163, 358)
    assert len(g) == 4 and len(s) == 3 and len(t) == 28
    kats = {
        "field_modulus": decimals(f"{REF}/sponge/test.rs", 5, 8)[0],
        "grain_lfsr": {"source": "R/sponge/poseidon/grain_lfsr.rs:190-218", "args": [False, 255, 3, 8, 31],
                       "rejection_sampling": g[:2], "mod_p": g[2:]},
        "sponge": {"source": "R/sponge/poseidon/mod.rs:381-404", "rate": 2, "optimized_for_weights": False,
                   "absorb": ["0", "1", "2"], "squeeze3": s},
        "default_params": {"source": "R/sponge/poseidon/traits.rs:163-358",
                           "constraints": {str(r): {"ark00": t[2 * (r - 2)], "mds00": t[2 * (r - 2) + 1]} for r in range(2, 9)},
                           "weights": {str(r): {"ark00": t[14 + 2 * (r - 2)], "mds00": t[14 + 2 * (r - 2) + 1]} for r in range(2, 9)}},
        "multiproof_prefix_lengths_8_leaves": {"source": "R/merkle_tree/tests/mod.rs:166", "value": [0, 2, 1, 2, 0, 2, 1, 2]},
    }
    json.dump(kats, open(f"{OUT}/reference_kats.json", "w"), indent=1)


if __name__ == "__main__":
    main()
