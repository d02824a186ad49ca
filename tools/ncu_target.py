"""Short workload for ncu captures: a few launches of the Poseidon CRH kernel (2^20 two-to-one hashes,
BLS12-381 Fr), or one 2^20-leaf Merkle build with `merkle`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N
from crypto_primitives_b200.crh.poseidon import TwoToOneCRH

cfg = cp.get_default_poseidon_parameters(cp.BLS12_381_FR, 2, False)
dev = torch.device("cuda:0")
n = 1 << 20
x = torch.randint(0, 2**59, (n, 2, 4), dtype=torch.int64).to(dev)
if len(sys.argv) > 1 and sys.argv[1] == "merkle":
    ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
    nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
    ctx = cfg.context(0)
    for _ in range(2):
        N.check(N.lib.cpb_merkle_poseidon_build_dev(ctx, ctx, x.data_ptr(), 2, n, ln.data_ptr(), nn.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
else:
    out = torch.empty((n, 4), dtype=torch.int64, device=dev)
    for _ in range(4):
        TwoToOneCRH.compress_batch_dev(cfg, x, out)
torch.cuda.synchronize()
