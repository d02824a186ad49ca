import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import crypto_primitives_b200 as cp
from crypto_primitives_b200.crh.poseidon import CRH, permute_batch
from helpers import synth_elems
from oracle import cref, fields as OF, poseidon as OP
p = OF.BLS12_381_FR
rng = OF.SplitMix64(3)
for (rate, rf, rp, alpha) in ((4, 2, 0, 5), (4, 2, 1, 5), (4, 8, 8, 5), (4, 8, 40, 5), (4, 8, 56, 5), (2, 8, 60, 5), (2, 8, 100, 5), (2, 8, 120, 5), (2, 8, 160, 5), (2, 8, 250, 5)):
    t = rate + 1
    ocfg = OP.PoseidonConfig(p, rf, rp, alpha, [[rng.field(p) for _ in range(t)] for _ in range(rf + rp)], [[rng.field(p) for _ in range(t)] for _ in range(t)], rate, 1)
    cfg = cp.PoseidonConfig.from_ints(cp.BLS12_381_FR, rf, rp, alpha, ocfg.mds, ocfg.ark, rate, 1)
    O = cref.Poseidon(ocfg)
    st = synth_elems(2100, (33, t), p)
    try:
        got = permute_batch(cfg, st)
    except Exception as e:
        print("t", t, "rf", rf, "rp", rp, "ERROR", e); continue
    exp = np.stack([O.permute(s) for s in st])
    nel = rf * t + 2 * t * t + t + max(rp, 1) + max(rp, 1) * (2 * t - 1) + rp * t + 1
    print("t", t, "rf", rf, "rp", rp, "consts bytes", nel * 32, "mismatch rows", int((got != exp).any(axis=(1, 2)).sum()), "of 33", flush=True)
