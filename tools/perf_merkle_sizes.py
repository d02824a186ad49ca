"""Development probe: device-resident cpb_merkle_poseidon_build_dev time for several tree sizes (the per-rank trees of
the 1/2/4/8-GPU runs of the 2^24-leaf job are 2^24 .. 2^21 leaves).  Run with CPB_TEAM_MAX / CPB_MERKLE_STREAMS to explore."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "bn254"
cfg = bench.poseidon_params(cp, name)
ctx = cfg.context(0)
for logn in (24, 23, 22, 21, 20, 16):
    n = 1 << logn
    leaves = torch.randint(0, 2**59, (n, 2, 4), dtype=torch.int64).to(dev)
    ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
    nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
    f = lambda: N.check(N.lib.cpb_merkle_poseidon_build_dev(ctx, ctx, leaves.data_ptr(), 2, n, ln.data_ptr(), nn.data_ptr(), st))
    for _ in range(2):
        f()
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ideal = (2 * n - 1) / (1.2697e8 if name == "bn254" else 1.4772e8) * 1e3
    print(f"{name} 2^{logn}: median {ts[2]:.3f} ms  min {ts[0]:.3f}  (bulk-rate ideal {ideal:.3f}, overhead {ts[2] - ideal:+.3f})  TEAM_MAX={os.environ.get('CPB_TEAM_MAX', 'default')} STREAMS={os.environ.get('CPB_MERKLE_STREAMS', 'default')}", flush=True)
    del leaves, ln, nn
