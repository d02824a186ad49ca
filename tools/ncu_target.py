"""Short workload for ncu captures.  usage: ncu_target.py [bls|bn254] [compress|merkle|top|pedersen|pedersen8] [log2 n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N
from crypto_primitives_b200.crh.poseidon import TwoToOneCRH

field = sys.argv[1] if len(sys.argv) > 1 else "bls"
what = sys.argv[2] if len(sys.argv) > 2 else "compress"
logn = int(sys.argv[3]) if len(sys.argv) > 3 else 20
if field == "bls":
    cfg = cp.get_default_poseidon_parameters(cp.BLS12_381_FR, 2, False)
else:
    ark, mds = cp.find_poseidon_ark_and_mds(cp.BN254_FR, 254, 2, 8, 57, 0)
    cfg = cp.PoseidonConfig(cp.BN254_FR, 8, 57, 5, mds, ark, 2, 1)
dev = torch.device("cuda:0")
n = 1 << logn
st = torch.cuda.current_stream().cuda_stream
if what in ("pedersen", "pedersen8"):
    from crypto_primitives_b200.commitment.pedersen import Commitment
    from crypto_primitives_b200.crh.pedersen import Window

    class Rng:
        def __init__(self, seed): self.g = np.random.default_rng(seed)
        def field(self, q): return int.from_bytes(self.g.bytes(40), "little") % q
    prm = Commitment.setup(Rng(1), Window(4, 256))
    if what == "pedersen8":
        prm.chunk_bits = 8                      # the shared-memory / TMA double-buffered kernel
    ctx = prm.context(0)
    inp = torch.randint(0, 256, (n, 128), dtype=torch.uint8, device=dev)
    out = torch.empty((n, 2, 4), dtype=torch.int64, device=dev)
    for _ in range(3):
        N.check(N.lib.cpb_pedersen_crh_batch_dev(ctx, inp.data_ptr(), 128, 128, out.data_ptr(), n, st))
else:
    x = torch.randint(0, 2**59, (n, 2, 4), dtype=torch.int64).to(dev)
    if what == "top":                               # the tree-top kernel: every level of a 2^logn-digest tree in one launch
        d = x[:, 0, :].contiguous()
        nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
        ctx = cfg.context(0)
        for _ in range(3):
            N.check(N.lib.cpb_merkle_poseidon_from_digests_dev(ctx, d.data_ptr(), n, nn.data_ptr(), st))
    elif what == "merkle":
        ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
        nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
        ctx = cfg.context(0)
        for _ in range(2):
            N.check(N.lib.cpb_merkle_poseidon_build_dev(ctx, ctx, x.data_ptr(), 2, n, ln.data_ptr(), nn.data_ptr(), st))
    else:
        out = torch.empty((n, 4), dtype=torch.int64, device=dev)
        for _ in range(3):
            TwoToOneCRH.compress_batch_dev(cfg, x, out)
torch.cuda.synchronize()
