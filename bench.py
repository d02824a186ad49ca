#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (driver contract; see DESIGN.md "Measurement").

Workload (BASELINE.json metric "Poseidon perms/sec & 2^24-leaf Merkle build s at 1/2/4/8 B200"):
one step = one full MerkleTree::new of 2^24 two-element leaves with Poseidon CRH leaves and
Poseidon two-to-one nodes over BN254 Fr (t=3, RF=8, RP=57, alpha=5) -- BASELINE.json configs[3].
It fits one GPU, so the same workload runs at N = 1, 2, 4, 8 (strong scaling): the leaves are
sharded contiguously over the ranks, each rank builds its subtree, ONE all-gather of the subtree
roots, top levels replicated (crypto_primitives_b200/distributed.py).
`value` = Poseidon permutations per second of the whole job = (2N-1) / step time, inputs resident in
HBM.  `e2e` = the same through the host-pointer C-ABI call (cpb_merkle_poseidon_build): leaves in
pinned host memory, H2D of the leaves and D2H of both node arrays inside the timed region.

--impl reference : the C restatement of the reference CPU path (oracle/cref, all host threads) on a
bounded sample of the same workload (the reference itself is Rust and cannot be built in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (field, log2 leaves, leaf_len, description)
    "merkle_2^24_poseidon_bn254": ("bn254", 24, 2, "2^24-leaf Poseidon Merkle tree, BN254 Fr (t=3, RF=8, RP=57, alpha=5), 2-element leaves"),
    "merkle_2^20_poseidon_bls12_381": ("bls", 20, 2, "2^20-leaf Poseidon Merkle tree, BLS12-381 Fr default rate-2 (alpha=17, RF=8, RP=31)"),
}
# parity-test configurations that can also be timed (single GPU only; not the contract line)
EXTRA_WORKLOADS = {
    "pedersen_crh_2^20_jubjub": "BASELINE config 3: 2^20 x 128-byte Pedersen CRH + commitment, Jubjub, window 4x256",
    "mixed_merkle_2^22": "BASELINE config 5 shape on one GPU: 2^22 x 128-byte leaves, Pedersen leaf hash (x-coordinate) + Poseidon two-to-one over BLS12-381 Fr",
}
DEFAULT_WORKLOAD = "merkle_2^24_poseidon_bn254"
HBM_PEAK_FALLBACK = 6650.0      # GB/s, B200_PROFILING.md fallback


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# --------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        load = [x for x in sm if x > 500] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------- parameters
def poseidon_params(cp, field_key):
    if field_key == "bls":
        return cp.get_default_poseidon_parameters(cp.BLS12_381_FR, 2, False)
    f = cp.BN254_FR
    ark, mds = cp.find_poseidon_ark_and_mds(f, 254, 2, 8, 57, 0)
    return cp.PoseidonConfig(f, 8, 57, 5, mds, ark, 2, 1)


def oracle_poseidon(field_key):
    from oracle import cref, fields as OF, poseidon as OP
    if field_key == "bls":
        cfg = OP.get_default_poseidon_parameters(OF.BLS12_381_FR, 2, False)
    else:
        ark, mds = OP.find_poseidon_ark_and_mds(OF.BN254_FR, 254, 2, 8, 57, 0)
        cfg = OP.PoseidonConfig(OF.BN254_FR, 8, 57, 5, ark, mds, 2, 1)
    return cfg, cref.Poseidon(cfg)


def synthetic_leaves_torch(torch, n, leaf_len, modulus, seed, device, pin=False):
    """n x leaf_len elements, each a uniformly random 4-limb value below the modulus' top limb (hence
    < p: a valid fully-reduced Montgomery representation).  Synthetic: there is no dataset for this path."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    top = (modulus >> 192) - 1
    x = torch.empty((n, leaf_len, 4), dtype=torch.int64, pin_memory=pin)
    chunk = 1 << 20
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        lo = torch.randint(-(2**63), 2**63 - 1, (e - s, leaf_len, 3), dtype=torch.int64, generator=g)
        hi = torch.randint(0, top, (e - s, leaf_len, 1), dtype=torch.int64, generator=g)
        x[s:e] = torch.cat([lo, hi], dim=2)
    return x.to(device) if device is not None else x


def merkle_launches(n: int) -> int:
    """Kernel launches of one cpb_merkle_poseidon_build_dev over n leaves (csrc/cpb_poseidon.cu: merkle_build_streams):
    S subtrees on S streams (leaf hash + their levels each), then the top log2 S levels."""
    h = n.bit_length() - 1
    S = 8
    while S > 1 and n // S < (1 << 14):
        S >>= 1
    lg = S.bit_length() - 1
    return S * (1 + (h - lg)) + lg if S > 1 else 1 + h


# --------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(field_key, leaf_len, threads, log_sample):
    """C restatement of the reference CPU path (kind "port") on a bounded sample: a 2^log_sample-leaf tree."""
    import numpy as np
    from oracle import cref
    cfg, P = oracle_poseidon(field_key)
    n = 1 << log_sample
    leaves = cref.synth_field_mont(0xB2000004, n * leaf_len, cfg.p).reshape(n, leaf_len, 4)
    t0 = time.perf_counter()
    cref.poseidon_merkle(P, P, leaves, threads=threads)
    dt = time.perf_counter() - t0
    perms = 2 * n - 1
    return perms / dt, dt, perms


def config1_probe():
    """BASELINE configs[0] -- the reference's own CPU-runnable case: crh::poseidon::CRH::evaluate on 1024 inputs of two
    BLS12-381 Fr elements (default rate-2 parameters).  CPU: the C restatement on ONE thread, as that config is stated;
    GPU: the host-pointer C-ABI call (copies included), median of 20.  Outputs compared."""
    import ctypes as C
    import numpy as np
    import crypto_primitives_b200 as cp
    from crypto_primitives_b200 import _native as N
    from oracle import cref
    ocfg, P = oracle_poseidon("bls")
    x = cref.synth_field_mont(0xB2000001, 1024 * 2, ocfg.p).reshape(1024, 2, 4)
    t0 = time.perf_counter()
    exp = P.crh_batch(x, threads=1)
    cpu_dt = time.perf_counter() - t0
    cfg = cp.get_default_poseidon_parameters(cp.BLS12_381_FR, 2, False)
    ctx = cfg.context(0)
    out = np.empty((1024, 4), dtype=np.uint64)
    times = []
    for _ in range(23):
        t0 = time.perf_counter()
        N.check(N.lib.cpb_poseidon_crh_batch(ctx, x.ctypes.data_as(N.u64p), 2, out.ctypes.data_as(N.u64p), 1024))
        times.append(time.perf_counter() - t0)
    gpu_dt = statistics.median(times[3:])
    return {"workload": "crh::poseidon::CRH::evaluate, BLS12-381 Fr, 1024 inputs x 2 elements",
            "cpu_single_thread": {"hashes_per_s": 1024 / cpu_dt, "ms": 1e3 * cpu_dt, "kind": "port"},
            "gpu_host_call": {"hashes_per_s": 1024 / gpu_dt, "ms": 1e3 * gpu_dt, "api": "cpb_poseidon_crh_batch (pageable host pointers, copies included)"},
            "outputs_equal": bool(np.array_equal(out, exp))}


def run_reference(args):
    """--impl reference: the CPU arm, on rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    field_key, logn, leaf_len, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    log_sample = min(logn, 17)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dt, perms = cpu_baseline(field_key, leaf_len, threads, log_sample)
        if i >= args.warmup:
            vals.append(v); times.append(dt)
    value = statistics.median(vals)
    sample = f"2^{log_sample}-leaf tree of the same shape per step ({2 * (1 << log_sample) - 1} permutations), {threads} threads"
    line = {"impl": "reference", "metric": "poseidon_perms_per_sec", "value": value, "unit": "perms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery)", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "note": "C restatement of the reference CPU path (oracle/cref); the Rust reference cannot be built here"},
            "cpu_baseline": {"value": value, "unit": "perms/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "perms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import crypto_primitives_b200 as cp
    from crypto_primitives_b200 import _native as N
    from crypto_primitives_b200.distributed import CudaPoseidonBackend, sharded_merkle_build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    field_key, logn, leaf_len, desc = WORKLOADS[args.workload]
    n_total = 1 << logn
    assert n_total % world == 0
    n_local = n_total // world
    params = poseidon_params(cp, field_key)
    p = params.field.modulus
    backend = CudaPoseidonBackend(params, params, local_rank)
    leaves = synthetic_leaves_torch(torch, n_local, leaf_len, p, 0xB2000004 + rank, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return sharded_merkle_build(backend, leaves, gather="roots")

    perms_total = 2 * n_total - 1
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    times = []
    root = None
    for _ in range(args.steps):
        flush.zero_()                                   # evict L2 between timed iterations (not timed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tree = step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        root = tree.root.clone()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = perms_total / (ms_per_step * 1e-3)

    # ---- e2e: host-pointer C-ABI call on this rank's shard, pinned host buffers, copies inside the timed region
    h_leaves = synthetic_leaves_torch(torch, n_local, leaf_len, p, 0xB2000004 + rank, None, pin=True)
    h_leaf_nodes = torch.empty((n_local, 4), dtype=torch.int64, pin_memory=True)
    h_nodes = torch.empty((max(n_local - 1, 1), 4), dtype=torch.int64, pin_memory=True)
    ctx = params.context(local_rank)

    def e2e_step():
        N.check(N.lib.cpb_merkle_poseidon_build(ctx, ctx, N.C.cast(h_leaves.data_ptr(), N.u64p), leaf_len, n_local,
                                                N.C.cast(h_leaf_nodes.data_ptr(), N.u64p), N.C.cast(h_nodes.data_ptr(), N.u64p)))
        if world > 1:
            r = h_nodes[0].to(dev, non_blocking=False).reshape(1, 4)
            roots = torch.empty((world, 4), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(roots, r)
            return backend.from_digests(roots)[0].cpu()
        return h_nodes[0]

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_root = e2e_step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = perms_total / (float(dt.item()) / e2e_steps)
    same_root = bool(torch.equal(torch.as_tensor(e2e_root).cpu().reshape(-1), root.cpu().reshape(-1)))

    # ---- roofline of the dominant kernel (k_poseidon_crh, the leaf-hash launch), timed live with CUDA events
    out = torch.empty((n_local, 4), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    N.check(N.lib.cpb_poseidon_crh_batch_dev(ctx, leaves.data_ptr(), leaf_len, out.data_ptr(), n_local, st))
    torch.cuda.synchronize()
    ks = []
    for _ in range(3):
        flush.zero_()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        N.check(N.lib.cpb_poseidon_crh_batch_dev(ctx, leaves.data_ptr(), leaf_len, out.data_ptr(), n_local, st))
        k1.record()
        torch.cuda.synchronize()
        ks.append(k0.elapsed_time(k1))
    k_ms = statistics.median(ks)
    alg_bytes = n_local * (32 * leaf_len + 32)                      # 32*L read + 32 written per leaf hash
    peaks = measured_peaks()
    peak = peaks["hbm_gbs"] if peaks else HBM_PEAK_FALLBACK
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # DRAM traffic of this kernel from the committed ncu --set full capture (profiles/r1_ncu_crh_*.txt: 2^20-hash launch,
    # dram__bytes_read.sum + dram__bytes_write.sum), scaled per hash to this launch
    ncu_bytes_per_hash = {"bn254": (67.304192e6 + 9.151744e6) / (1 << 20), "bls": (67.3e6 + 9.6e6) / (1 << 20)}[field_key]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_bytes_per_hash * n_local, "traffic_source": "ncu capture of a 2^20-hash launch, scaled per hash",
                "kernel": "k_poseidon_crh (leaf level: %d hashes of %d elements)" % (n_local, leaf_len), "kernel_ms": k_ms,
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "B200_PROFILING.md fallback (of fallback)",
                "note": "the path is bound by the integer multiply pipe, not HBM (~6e4 IMAD-class instructions per 96 algorithmic bytes); see integer_pipe"}
    # integer-pipe view: wide 32x32->64 multiply-adds needed by the schedule vs the measured issue rate
    wide_per_perm = {"bn254": 61896, "bls": 44784}[field_key]           # DESIGN.md §4.2; matches the ncu opcode mix
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    int_peak = 148 * 32 * sm_clock * 1e6                                # IMAD.WIDE/IMAD.HI: 32 lanes/clk/SM (tools/ubench_int.cu, ncu)
    integer = {"wide_madds_per_perm": wide_per_perm, "achieved_wide_madds_per_s": n_local * wide_per_perm / (k_ms * 1e-3),
               "peak_wide_madds_per_s": int_peak, "frac": n_local * wide_per_perm / (k_ms * 1e-3) / int_peak,
               "peak_source": "148 SMs x 32 lanes/clk (measured IMAD.WIDE rate) x sampled SM clock"}

    if rank == 0:
        launches_per_step = merkle_launches(n_local) + ((world.bit_length() - 1) if world > 1 else 0)
        threads = os.cpu_count() or 1
        cpu_v, cpu_dt, cpu_perms = cpu_baseline(field_key, leaf_len, threads, min(logn, 18))
        line = {"metric": "poseidon_perms_per_sec", "value": value, "unit": "perms/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "merkle_build_s": ms_per_step * 1e-3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integer)", "data": "synthetic",
                "config": {"workload": args.workload, "description": desc, "leaves_total": n_total, "leaves_per_gpu": n_local,
                           "parallelism": f"leaf-sharded x{world}, one all-gather of subtree roots", "l2": "flushed (256 MB write) between timed steps",
                           "perms_per_step": perms_total},
                "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
                "e2e": {"value": e2e_value, "unit": "perms/s", "h2d_bytes_per_step": n_local * leaf_len * 32 * world,
                        "d2h_bytes_per_step": (2 * n_local - 1) * 32 * world, "steps": e2e_steps, "root_matches_device_run": same_root,
                        "api": "cpb_merkle_poseidon_build (host pointers, pinned)"},
                "roofline": roofline, "integer_pipe": integer,
                "cpu_baseline": {"value": cpu_v, "unit": "perms/s", "cores": threads, "kind": "port",
                                 "sample": f"2^{min(logn, 18)}-leaf tree of the same shape ({cpu_perms} permutations, {cpu_dt:.2f} s wall)"}}
        if world == 1:
            try:
                line["config1"] = config1_probe()
            except Exception as e:                      # an extra, never a reason to lose the contract line
                line["config1"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_extra(args):
    """Device-resident timing of the Pedersen / mixed-tree configurations (one GPU)."""
    import numpy as np
    import torch
    import crypto_primitives_b200 as cp
    from crypto_primitives_b200 import _native as N
    from crypto_primitives_b200.commitment.pedersen import Commitment
    from crypto_primitives_b200.crh.pedersen import Window

    class Rng:
        def __init__(self, seed): self.g = np.random.default_rng(seed)
        def field(self, q): return int.from_bytes(self.g.bytes(40), "little") % q

    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    prm = Commitment.setup(Rng(0xB2000003), Window(4, 256))
    ctx = prm.context(0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        ts = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    g = torch.Generator(device="cpu").manual_seed(3)
    if args.workload == "pedersen_crh_2^20_jubjub":
        n = 1 << 20
        inp = torch.randint(0, 256, (n, 128), dtype=torch.uint8, generator=g).to(dev)
        rnd = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
        rnd[:, 31] &= 0x0F
        rnd = rnd.to(dev)
        out = torch.empty((n, 2, 4), dtype=torch.int64, device=dev)
        ms_crh = timed(lambda: N.check(N.lib.cpb_pedersen_crh_batch_dev(ctx, inp.data_ptr(), 128, 128, out.data_ptr(), n, st)))
        ms_com = timed(lambda: N.check(N.lib.cpb_pedersen_commit_batch_dev(ctx, inp.data_ptr(), 128, 128, rnd.data_ptr(), out.data_ptr(), n, st)))
        line = {"metric": "pedersen_hashes_per_sec", "value": n / (ms_crh * 1e-3), "unit": "hashes/s", "ms_per_step": ms_crh,
                "commit_per_sec": n / (ms_com * 1e-3), "commit_ms": ms_com}
    else:
        n = 1 << 22
        node = poseidon_params(cp, "bls")
        nctx = node.context(0)
        leaves = torch.randint(0, 256, (n, 128), dtype=torch.uint8, generator=g).to(dev)
        ln = torch.empty((n, 4), dtype=torch.int64, device=dev)
        nn = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
        ms = timed(lambda: N.check(N.lib.cpb_merkle_mixed_build_dev(ctx, nctx, leaves.data_ptr(), 128, 128, n, ln.data_ptr(), nn.data_ptr(), st)))
        line = {"metric": "mixed_merkle_build_s", "value": ms * 1e-3, "unit": "s", "ms_per_step": ms, "higher_is_better": False,
                "hashes_per_step": {"pedersen": n, "poseidon": n - 1}}
    line.update({"n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "data": "synthetic", "dtype": "u32x8 (256-bit Montgomery integer)",
                 "config": {"workload": args.workload, "description": EXTRA_WORKLOADS[args.workload], "l2": "flushed between timed steps"}})
    line.setdefault("higher_is_better", True)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS) + list(EXTRA_WORKLOADS))
    args = ap.parse_args()
    if args.workload in EXTRA_WORKLOADS:
        if args.impl != "b200" or args.gpus != 1:
            raise SystemExit("the extra workloads are single-GPU, --impl b200 only")
        run_extra(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
