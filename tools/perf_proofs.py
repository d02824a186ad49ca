"""Development probe in the shape of the reference's merkle_tree bench (benches/merkle_tree.rs:36-209: 2^20 leaves --
create / prove all / verify all / multiproof), here with the Poseidon field-leaf Config over BLS12-381 Fr.
Wall-clock through the host-pointer API (copies included); the proofs are generated as arrays by index arithmetic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench, bench_inputs as BI
import crypto_primitives_b200 as cp
from crypto_primitives_b200 import _native as N
from crypto_primitives_b200.merkle_tree import MerkleTree


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << logn
    cfg = bench.poseidon_params(cp, "bls")
    leaves = BI.field_elements_torch(torch, N, cfg.field.id, BI.SEED_CONFIG2, 0, 2 * n, 0).view(n, 2, 4).cpu().numpy().view(np.uint64)
    MerkleTree.new(cfg, cfg, leaves[:1024])                                   # warm-up: context, kernels
    t0 = time.perf_counter(); tree = MerkleTree.new(cfg, cfg, leaves); t_create = time.perf_counter() - t0
    t0 = time.perf_counter(); proofs = tree.generate_proofs_batch(np.arange(n)); t_prove = time.perf_counter() - t0
    tree.verify_proofs_batch(tuple(a[:1024] for a in proofs), leaves[:1024])
    t0 = time.perf_counter(); ok = tree.verify_proofs_batch(proofs, leaves); t_verify = time.perf_counter() - t0
    assert ok.all()
    k = 1 << 12
    sel = np.sort(np.random.default_rng(1).choice(n, k, replace=False))
    t0 = time.perf_counter(); mp = tree.generate_multi_proof(sel); t_mp = time.perf_counter() - t0
    t0 = time.perf_counter(); okm = mp.verify(cfg, cfg, tree.root(), leaves[sel]); t_mpv = time.perf_counter() - t0
    assert okm
    upd = np.sort(np.random.default_rng(2).choice(n, k, replace=False))
    t0 = time.perf_counter(); tree.update_batch(upd, leaves[(upd + 1) % n]); t_upd = time.perf_counter() - t0
    print(f"poseidon BLS12-381 Fr tree, 2^{logn} leaves x 2 elements (host API, copies included):")
    print(f"  create            {1e3 * t_create:9.2f} ms")
    print(f"  prove all         {1e3 * t_prove:9.2f} ms  (index arithmetic on the host arrays, {proofs[1].nbytes / 2**20:.0f} MiB of paths)")
    print(f"  verify all        {1e3 * t_verify:9.2f} ms  ({n * (logn + 1) / t_verify / 1e6:.1f} M permutations/s incl. H2D of the paths; one launch)")
    print(f"  multiproof gen    {1e3 * t_mp:9.2f} ms  ({k} leaves)")
    print(f"  multiproof verify {1e3 * t_mpv:9.2f} ms  ({k} leaves, {logn} level-synchronous batches)")
    print(f"  update {k} leaves {1e3 * t_upd:9.2f} ms  ({logn} level-synchronous batches)")


if __name__ == "__main__":
    main()
