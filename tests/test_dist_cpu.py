"""world_size-2 and -4 gloo tests (CPU) of the leaf-sharded Merkle build: the sharding, the single
all-gather of subtree roots, the replicated top levels and the per-level all-gather assembly.  The
hashing backend is the C oracle on CPU tensors (this is a test of the host logic around the kernels;
the CUDA backend is exercised by tests/test_gpu_dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    digest_words = 4

    def __init__(self):
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle_config
        from oracle import cref
        self.cref = cref
        self.P = cref.Poseidon(oracle_config("bls_default_r2")[1])

    def _np(self, t):
        return t.numpy().view(np.uint64)

    def build_local(self, leaves):
        ln, nn = self.cref.poseidon_merkle(self.P, self.P, self._np(leaves), threads=1)
        return torch.from_numpy(ln.view(np.int64)), torch.from_numpy(nn.view(np.int64))

    def hash_leaves(self, leaves):
        return torch.from_numpy(self.P.crh_batch(self._np(leaves)).view(np.int64))

    def from_digests(self, d):
        lvl = self._np(d.contiguous())
        n = lvl.shape[0]
        nodes = np.empty((n - 1, 4), dtype=np.uint64)
        start = n // 2 - 1
        nodes[start:] = self.P.compress_batch(lvl.reshape(-1, 2, 4))
        while start > 0:
            upper, start = start, (start - 1) // 2
            nodes[start:upper] = self.P.compress_batch(nodes[upper:2 * upper + 1].reshape(-1, 2, 4))
        return torch.from_numpy(nodes.view(np.int64))


def _worker(rank, world, port, n, gather, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import importlib.util
        spec = importlib.util.spec_from_file_location("cpb_dist", os.path.join(ROOT, "crypto_primitives_b200", "distributed.py"))
        D = importlib.util.module_from_spec(spec)
        sys.modules["cpb_dist"] = D
        spec.loader.exec_module(D)
        from helpers import oracle_config, synth_elems
        be = OracleBackend()
        p = oracle_config("bls_default_r2")[1].p
        leaves = synth_elems(55, (n, 2), p)
        per = n // world
        local = torch.from_numpy(np.ascontiguousarray(leaves[rank * per:(rank + 1) * per]).view(np.int64))
        tree = D.sharded_merkle_build(be, local, gather=gather)
        exp_leaf, exp_nodes = be.cref.poseidon_merkle(be.P, be.P, leaves, threads=1)
        ok = np.array_equal(tree.root.numpy().view(np.uint64), exp_nodes[0]) and tree.height() == n.bit_length()
        ok &= np.array_equal(tree.local_leaf_nodes.numpy().view(np.uint64), exp_leaf[rank * per:(rank + 1) * per])
        if world > 1:
            ok &= np.array_equal(tree.top_nodes.numpy().view(np.uint64), exp_nodes[: world - 1])
        for gstart, cnt, lstart in D.level_slices(n, world, rank):
            ok &= np.array_equal(tree.local_nodes[lstart:lstart + cnt].numpy().view(np.uint64), exp_nodes[gstart:gstart + cnt])
        if gather == "levels":
            ok &= np.array_equal(tree.non_leaf_nodes.numpy().view(np.uint64), exp_nodes)
            ok &= np.array_equal(tree.leaf_nodes.numpy().view(np.uint64), exp_leaf)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,n,gather", [(2, 64, "roots"), (2, 64, "levels"), (4, 32, "levels"), (2, 2, "roots")])
def test_sharded_build_gloo(world, n, gather):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, gather, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True) for r in range(world)]
    assert all(p.exitcode == 0 for p in procs)
