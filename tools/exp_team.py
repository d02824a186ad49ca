"""Development probe: 2^20- and 2^16-leaf tree build time vs CPB_TEAM_MAX / CPB_MERKLE_STREAMS (set in the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crypto_primitives_b200 import _native as N
from tools.quick_perf import cfg_for, timeit

dev = torch.device("cuda:0")
for name in ("bls", "bn254"):
    cfg = cfg_for(name)
    ctx = cfg.context(0)
    for logn in (16, 20):
        n = 1 << logn
        leaves = torch.randint(0, 2**59, (n, 2, 4), dtype=torch.int64).to(dev)
        ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
        nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        def build():
            N.check(N.lib.cpb_merkle_poseidon_build_dev(ctx, ctx, leaves.data_ptr(), 2, n, ln.data_ptr(), nn.data_ptr(), st))
        ms = timeit(build, iters=10, warm=3)
        print(f"TEAM_MAX={os.environ.get('CPB_TEAM_MAX','default')} STREAMS={os.environ.get('CPB_MERKLE_STREAMS','default')} {name} merkle 2^{logn}: {ms:.3f} ms", flush=True)
