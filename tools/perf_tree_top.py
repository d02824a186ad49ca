"""Development probe: latency of the small levels of a tree (the tree-top kernel) -- cpb_merkle_poseidon_from_digests_dev
for n = 2 .. 2^16 digests, and a flat two-to-one batch of 32 .. 4096 hashes (one level)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    for name in ("bn254", "bls"):
        cfg = bench.poseidon_params(cp, name)
        ctx = cfg.context(0)
        for logn in (1, 2, 3, 6, 10, 13, 14, 16):
            n = 1 << logn
            d = torch.randint(0, 2**59, (n, 4), dtype=torch.int64).to(dev)
            nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
            ms = timeit(lambda: N.check(N.lib.cpb_merkle_poseidon_from_digests_dev(ctx, d.data_ptr(), n, nn.data_ptr(), st)))
            print(f"{name} from_digests n=2^{logn}: {ms:.4f} ms  ({ms / logn:.4f} ms/level)", flush=True)
        for m in (32, 1024, 4096):
            pairs = torch.randint(0, 2**59, (m, 2, 4), dtype=torch.int64).to(dev)
            out = torch.empty((m, 4), dtype=torch.int64, device=dev)
            ms = timeit(lambda: N.check(N.lib.cpb_poseidon_compress_batch_dev(ctx, pairs.data_ptr(), out.data_ptr(), m, st)))
            print(f"{name} compress {m} hashes: {ms:.4f} ms", flush=True)


if __name__ == "__main__":
    main()
