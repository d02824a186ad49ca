"""Development perf probe: device-resident throughput of the Pedersen kernels (Jubjub, window 4x256)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N
from crypto_primitives_b200.commitment.pedersen import Commitment
from crypto_primitives_b200.crh.pedersen import Window


class Rng:
    def __init__(self, seed): self.g = np.random.default_rng(seed)
    def field(self, q): return int.from_bytes(self.g.bytes(40), "little") % q


def timeit(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


import time
from crypto_primitives_b200.crh.pedersen import Parameters
base = Commitment.setup(Rng(1), Window(4, 256))
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for cb in tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (8, 12, 16):
  prm = Parameters(base.curve, base.window, base.generators, base.randomness_generator, chunk_bits=cb)
  t0 = time.time(); ctx = prm.context(0); print(f"chunk_bits={cb}: context (table build) {time.time() - t0:.3f} s", flush=True)
  for logn in (20,):
    n = 1 << logn
    inp = torch.randint(0, 256, (n, 128), dtype=torch.uint8, device=dev)
    rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
    rnd[:, 31] &= 0x0F
    out = torch.empty((n, 2, 4), dtype=torch.int64, device=dev)
    ms = timeit(lambda: N.check(N.lib.cpb_pedersen_crh_batch_dev(ctx, inp.data_ptr(), 128, 128, out.data_ptr(), n, st)))
    print(f"  pedersen crh n=2^{logn}: {ms:.3f} ms  {n / ms / 1e3:.2f} M hashes/s", flush=True)
    ms = timeit(lambda: N.check(N.lib.cpb_pedersen_commit_batch_dev(ctx, inp.data_ptr(), 128, 128, rnd.data_ptr(), out.data_ptr(), n, st)))
    print(f"  pedersen commit n=2^{logn}: {ms:.3f} ms  {n / ms / 1e3:.2f} M commits/s", flush=True)
