"""Multi-GPU test of the leaf-sharded Merkle build with the CUDA backend over NCCL (needs >= 2 GPUs;
skipped on a single-GPU box -- the sharding logic itself is covered on CPU by tests/test_dist_cpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, logn, gather, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle_config, product_config, synth_elems
        from crypto_primitives_b200.distributed import CudaPoseidonBackend, level_slices, sharded_merkle_build
        from oracle import cref
        _, ocfg = oracle_config("bn254_r2")
        cfg = product_config("bn254_r2")
        n = 1 << logn
        leaves = synth_elems(77, (n, 2), ocfg.p)
        per = n // world
        local = torch.from_numpy(np.ascontiguousarray(leaves[rank * per:(rank + 1) * per]).view(np.int64)).cuda()
        tree = sharded_merkle_build(CudaPoseidonBackend(cfg, cfg, rank), local, gather=gather)
        torch.cuda.synchronize()
        O = cref.Poseidon(ocfg)
        exp_leaf, exp_nodes = cref.poseidon_merkle(O, O, leaves, threads=8)
        ok = np.array_equal(tree.root.cpu().numpy().view(np.uint64), exp_nodes[0])
        for gstart, cnt, lstart in level_slices(n, world, rank):
            ok &= np.array_equal(tree.local_nodes[lstart:lstart + cnt].cpu().numpy().view(np.uint64), exp_nodes[gstart:gstart + cnt])
        if gather == "levels":
            ok &= np.array_equal(tree.non_leaf_nodes.cpu().numpy().view(np.uint64), exp_nodes)
            ok &= np.array_equal(tree.leaf_nodes.cpu().numpy().view(np.uint64), exp_leaf)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gather", ["roots", "levels"])
def test_sharded_build_nccl(gather):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 14, gather, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True) for r in range(world)]


def _worker_mixed(rank, world, port, logn, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import crypto_primitives_b200 as cp
        from helpers import oracle_config, product_config
        from crypto_primitives_b200.crh.pedersen import Parameters, Window
        from crypto_primitives_b200.distributed import CudaMixedBackend, sharded_merkle_build
        from oracle import cref, pedersen as OPD
        ow = OPD.Window(4, 256)
        oprm = OPD.setup(ow, 5)
        g = cp.BLS12_381_FR.elements([c for w in oprm.generators for pt in w for c in pt]).reshape(256, 4, 2, 4)
        prm = Parameters(cp.curves.JUBJUB, Window(4, 256), g)
        _, ocfg = oracle_config("bls_default_r2")
        node = product_config("bls_default_r2")
        n = 1 << logn
        leaves = np.ascontiguousarray(cref.synth_bytes(31, n * 128).reshape(n, 128))
        per = n // world
        local = torch.from_numpy(leaves[rank * per:(rank + 1) * per].copy()).cuda()
        tree = sharded_merkle_build(CudaMixedBackend(prm, node, rank), local, gather="levels")
        torch.cuda.synchronize()
        exp_leaf, exp_nodes = cref.mixed_merkle(cref.Pedersen(oprm, ow), cref.Poseidon(ocfg), leaves, threads=8)
        ok = np.array_equal(tree.non_leaf_nodes.cpu().numpy().view(np.uint64), exp_nodes)
        ok &= np.array_equal(tree.leaf_nodes.cpu().numpy().view(np.uint64), exp_leaf)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_mixed_tree_nccl():
    """BASELINE config 5 shape (Pedersen leaf hash + Poseidon two-to-one) sharded over the available GPUs."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_mixed, args=(r, world, port, 10, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert sorted(q.get(timeout=5) for _ in range(world)) == [(r, True) for r in range(world)]
