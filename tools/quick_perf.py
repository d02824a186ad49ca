"""Development perf probe (not the contract bench): device-resident throughput of the Poseidon kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N
from crypto_primitives_b200.crh.poseidon import CRH, TwoToOneCRH


def cfg_for(name):
    if name == "bls":
        return cp.get_default_poseidon_parameters(cp.BLS12_381_FR, 2, False)
    f = cp.BN254_FR
    ark, mds = cp.find_poseidon_ark_and_mds(f, 254, 2, 8, 57, 0)
    return cp.PoseidonConfig(f, 8, 57, 5, mds, ark, 2, 1)


def timeit(fn, iters=5, warm=2):
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
    for name in ("bls", "bn254"):
        cfg = cfg_for(name)
        p = cfg.field.modulus
        for logn in (16, 20, 22):
            n = 1 << logn
            g = torch.Generator(device="cpu").manual_seed(1)
            raw = torch.randint(0, 2**62, (n, 2, 4), dtype=torch.int64, generator=g)
            raw[..., 3] &= (1 << 59) - 1          # < 2^251 < p: canonical-looking limbs (values need not be meaningful)
            x = raw.to(dev)
            out = torch.empty((n, 4), dtype=torch.int64, device=dev)
            ms = timeit(lambda: TwoToOneCRH.compress_batch_dev(cfg, x, out))
            print(f"{name} compress n=2^{logn}: {ms:.3f} ms  {n / ms / 1e3:.2f} M perms/s", flush=True)
        # merkle build 2^20
        n = 1 << 20
        leaves = torch.randint(0, 2**59, (n, 2, 4), dtype=torch.int64).to(dev)
        ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
        nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
        ctx = cfg.context(0)
        st = torch.cuda.current_stream().cuda_stream
        def build():
            N.check(N.lib.cpb_merkle_poseidon_build_dev(ctx, ctx, leaves.data_ptr(), 2, n, ln.data_ptr(), nn.data_ptr(), st))
        ms = timeit(build, iters=3, warm=1)
        print(f"{name} merkle 2^20: {ms:.3f} ms  {(2 * n - 1) / ms / 1e3:.2f} M perms/s", flush=True)


if __name__ == "__main__":
    main()
