"""CUDA-event breakdown of one rank's step of the 8-GPU run (a 2^21-leaf BN254 tree): the leaf hash, every level as its own
full-width launch, the small levels in the tree-top kernel, and the whole build call for comparison -- where the time of a
sharded build goes and what is left of it when the bulk rate is subtracted."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 21
cfg = bench.poseidon_params(cp, "bn254")
ctx = cfg.context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


n = 1 << logn
rate = None
leaves = torch.randint(0, 2**59, (n, 2, 4), dtype=torch.int64).to(dev)
ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
rows = []
t_leaf = timed(lambda: N.check(N.lib.cpb_poseidon_crh_batch_dev(ctx, leaves.data_ptr(), 2, ln.data_ptr(), n, st)))
rate = n / t_leaf                                   # hashes per ms at the bulk rate
rows.append((f"leaf hash, 2^{logn} hashes", n, t_leaf))
total = t_leaf
for l in range(logn - 1, 12, -1):
    m = 1 << l
    src = leaves.view(-1, 4)[: 2 * m].view(m, 2, 4)
    t = timed(lambda: N.check(N.lib.cpb_poseidon_compress_batch_dev(ctx, src.data_ptr(), nn.data_ptr(), m, st)))
    rows.append((f"level of 2^{l} hashes, one launch", m, t))
    total += t
d = leaves.view(-1, 4)[: 1 << 13]
t_top = timed(lambda: N.check(N.lib.cpb_merkle_poseidon_from_digests_dev(ctx, d.data_ptr(), 1 << 13, nn.data_ptr(), st)))
rows.append(("13 levels 4096 .. 1 hashes, ONE tree-top launch", (1 << 13) - 1, t_top))
total += t_top
t_build = timed(lambda: N.check(N.lib.cpb_merkle_poseidon_build_dev(ctx, ctx, leaves.data_ptr(), 2, n, ln.data_ptr(), nn.data_ptr(), st)))
print(f"BN254 Fr, 2^{logn}-leaf tree on one GPU (= one rank of the {1 << (24 - logn)}-GPU run); bulk rate {rate / 1e3:.2f} M hashes/s from the leaf launch")
print(f"{'phase':55s} {'hashes':>9s} {'ms':>9s} {'at bulk rate':>13s} {'excess':>8s}")
for name, m, t in rows:
    print(f"{name:55s} {m:9d} {t:9.3f} {m / rate:13.3f} {t - m / rate:+8.3f}")
print(f"{'sum of the phases run one after the other':55s} {2 * n - 1:9d} {total:9.3f} {(2 * n - 1) / rate:13.3f} {total - (2 * n - 1) / rate:+8.3f}")
print(f"{'cpb_merkle_poseidon_build_dev (8 subtree streams + tree-tops + top)':55s} {2 * n - 1:9d} {t_build:9.3f} {(2 * n - 1) / rate:13.3f} {t_build - (2 * n - 1) / rate:+8.3f}")
