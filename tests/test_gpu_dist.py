"""Multi-GPU tests of the leaf-sharded Merkle build (need >= 2 GPUs; skipped on a single-GPU box -- the sharding logic
itself is covered on CPU by tests/test_dist_cpu.py, the fused root-exchange kernel on one GPU by
tests/test_gpu_merkle.py::test_fused_root_exchange_with_all_ranks_on_one_gpu): one process per GPU with the roots
exchanged by the fused peer-memory kernel (CUDA IPC over NVLink) or by an NCCL all-gather, and one process driving
all GPUs through cpb_merkle_poseidon_build_multi (peer memory and NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, logn, gather, q, fused=False):
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
        from crypto_primitives_b200.distributed import CudaPoseidonBackend, Exchange, level_slices, sharded_merkle_build
        from oracle import cref
        _, ocfg = oracle_config("bn254_r2")
        cfg = product_config("bn254_r2")
        n = 1 << logn
        leaves = synth_elems(77, (n, 2), ocfg.p)
        per = n // world
        local = torch.from_numpy(np.ascontiguousarray(leaves[rank * per:(rank + 1) * per]).view(np.int64)).cuda()
        ex = Exchange(rank) if fused else None
        be = CudaPoseidonBackend(cfg, cfg, rank)
        for _ in range(3 if fused else 1):                     # several collective calls: epochs / slot parity
            tree = sharded_merkle_build(be, local, gather=gather, exchange=ex)
        torch.cuda.synchronize()
        O = cref.Poseidon(ocfg)
        exp_leaf, exp_nodes = cref.poseidon_merkle(O, O, leaves, threads=8)
        ok = np.array_equal(tree.root.cpu().numpy().view(np.uint64), exp_nodes[0])
        ok &= np.array_equal(tree.top_nodes.cpu().numpy().view(np.uint64), exp_nodes[:world - 1])
        for gstart, cnt, lstart in level_slices(n, world, rank):
            ok &= np.array_equal(tree.local_nodes[lstart:lstart + cnt].cpu().numpy().view(np.uint64), exp_nodes[gstart:gstart + cnt])
        if gather == "levels":
            ok &= np.array_equal(tree.non_leaf_nodes.cpu().numpy().view(np.uint64), exp_nodes)
            ok &= np.array_equal(tree.leaf_nodes.cpu().numpy().view(np.uint64), exp_leaf)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("gather", ["roots", "levels"])
def test_sharded_build_nccl(gather, fused):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 14, gather, q, fused)) for r in range(world)]
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
        from crypto_primitives_b200.distributed import CudaMixedBackend, Exchange, sharded_merkle_build
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
        tree = sharded_merkle_build(CudaMixedBackend(prm, node, rank), local, gather="levels", exchange=Exchange(rank))
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


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_single_process_multi_gpu_build_through_the_c_abi(exchange):
    """cpb_merkle_poseidon_build_multi: ONE process, all GPUs, host arrays in the reference's layout; the root exchange
    through peer memory (fused kernel) and through ncclCommInitAll + ncclAllGather.  Run in a subprocess: the exchange
    mode is read from the environment when the group is created."""
    import subprocess
    import torch
    world = min(torch.cuda.device_count(), 8)
    world = 1 << (world.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    code = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from helpers import oracle_config, product_config, synth_elems
from crypto_primitives_b200 import _native as N
from oracle import cref
world = %d
_, ocfg = oracle_config("bn254_r2")
cfg = product_config("bn254_r2")
devs = (C.c_int * world)(*range(world))
m = N.vp()
N.check(N.lib.cpb_multi_create(world, devs, C.byref(m)))
assert N.lib.cpb_multi_uses_nccl(m) == (1 if %r == "nccl" else 0)
ctxs = (N.vp * world)(*[cfg.context(d) for d in range(world)])
O = cref.Poseidon(ocfg)
for logn in (2 + world.bit_length(), 13, 17):
    n = 1 << logn
    leaves = synth_elems(55 + logn, (n, 2), ocfg.p)
    ln = np.zeros((n, 4), dtype=np.uint64); nn = np.zeros((n - 1, 4), dtype=np.uint64)
    for rep in range(2):
        N.check(N.lib.cpb_merkle_poseidon_build_multi(m, ctxs, ctxs, leaves.ctypes.data_as(N.u64p), 2, n, ln.ctypes.data_as(N.u64p), nn.ctypes.data_as(N.u64p)))
    el, en = cref.poseidon_merkle(O, O, leaves, threads=8)
    assert np.array_equal(ln, el) and np.array_equal(nn, en), logn
N.lib.cpb_multi_destroy(m)
print("multi ok")
""" % (ROOT, os.path.join(ROOT, "tests"), world, exchange)
    env = dict(os.environ, CPB_MULTI_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "multi ok" in r.stdout, r.stdout + r.stderr


def test_cpp_program_builds_a_tree_on_all_gpus():
    """tests/cpp/test_multi.cpp: a plain C++ host linked against libcpb200.so (what a Rust shim's FFI does) builds one tree
    on one GPU and on all of them and compares the arrays."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    out_dir = os.path.join(ROOT, "tests", "host", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "test_multi")
    lib_dir = os.path.join(ROOT, "crypto_primitives_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_multi.cpp"),
                           "-L", lib_dir, "-l:libcpb200.so", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "multi-gpu build ok" in r.stdout, r.stdout + r.stderr
