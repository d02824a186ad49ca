"""Counter-based synthetic inputs shared by bench.py, tests/golden/make_bench_goldens.py and the tests
(SURVEY.md §8d): one SplitMix64 stream per workload, addressed by GLOBAL index, so every rank of an
N-GPU run generates exactly its slice of the same input and the result does not depend on N.

  stream(seed)[k]   = splitmix64 output number k+1 (k = 0, 1, ...): z = seed + (k+1)*0x9E3779B97F4A7C15, then the
                      two xor-shift-multiply rounds and the final xor-shift -- the generator of oracle/cref.py
                      (`splitmix64_stream`) and oracle/fields.py (`SplitMix64`).
  field element e   = stream[4e .. 4e+3] as little-endian limbs, reduced mod p, converted to Montgomery form
                      (== oracle.cref.synth_field_mont(seed, n, p)[e]).
  byte b            = byte b of the little-endian stream (== oracle.cref.synth_bytes(seed, n)[b]).

No oracle import here: the numpy versions restate the generator, the torch versions run it on the
GPU and convert with the library's own cpb_field_to_montgomery_dev.
"""
from __future__ import annotations

import numpy as np

GOLDEN_GAMMA = 0x9E3779B97F4A7C15
M1 = 0xBF58476D1CE4E5B9
M2 = 0x94D049BB133111EB
MASK64 = (1 << 64) - 1

# workload seeds (0xB2000000 + BASELINE config number; auxiliary streams +0x10)
SEED_CONFIG1 = 0xB2000001      # 1024 CRH inputs, BLS12-381 Fr
SEED_CONFIG2 = 0xB2000002      # 2^20-leaf Poseidon tree, BLS12-381 Fr
SEED_CONFIG2_PERM = 0xB2000012  # 2^22 bare permutation states, BLS12-381 Fr
SEED_CONFIG3 = 0xB2000003      # 2^20 x 128-byte Pedersen inputs
SEED_CONFIG3_RAND = 0xB2000013  # their commitment randomness
SEED_CONFIG3_PARAMS = 0xB2000023  # Pedersen generators (config 3 and 5)
SEED_CONFIG4 = 0xB2000004      # 2^24-leaf Poseidon tree, BN254 Fr
SEED_CONFIG5 = 0xB2000005      # 2^22 x 128-byte leaves of the mixed tree


def splitmix64_np(seed: int, start: int, count: int) -> np.ndarray:
    """stream(seed)[start : start+count] as uint64."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + idx * np.uint64(GOLDEN_GAMMA)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(M2)
        return z ^ (z >> np.uint64(31))


def raw_field_limbs_np(seed: int, elem_start: int, n_elems: int) -> np.ndarray:
    """(n_elems, 4) uint64 raw limbs (NOT reduced, NOT Montgomery) of elements elem_start..."""
    return splitmix64_np(seed, 4 * elem_start, 4 * n_elems).reshape(n_elems, 4)


def bytes_np(seed: int, byte_start: int, nbytes: int) -> np.ndarray:
    w0 = byte_start // 8
    w1 = (byte_start + nbytes + 7) // 8
    b = splitmix64_np(seed, w0, w1 - w0).view(np.uint8)
    off = byte_start - 8 * w0
    return b[off:off + nbytes].copy()


def randomness_np(seed: int, start: int, n: int) -> np.ndarray:
    """n x 32-byte little-endian scalars below 2^251 (< the Jubjub scalar modulus): commitment randomness."""
    r = bytes_np(seed, 32 * start, 32 * n).reshape(n, 32)
    r[:, 31] &= 0x07
    return r


def _s64(v: int) -> int:
    v &= MASK64
    return v - (1 << 64) if v >> 63 else v


def splitmix64_torch(torch, seed: int, start: int, count: int, device):
    """stream(seed)[start : start+count] as an int64 tensor on `device` (two's-complement image of the uint64 values).
    int64 arithmetic wraps; `>>` is arithmetic, so the sign-extended bits are masked off."""
    idx = torch.arange(start + 1, start + count + 1, dtype=torch.int64, device=device)
    z = idx * _s64(GOLDEN_GAMMA) + _s64(seed)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(M1)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(M2)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


def field_elements_torch(torch, native, field_id: int, seed: int, elem_start: int, n_elems: int, device_index: int, chunk: int = 1 << 24):
    """(n_elems, 4) int64 Montgomery limbs on cuda:device_index == oracle.cref.synth_field_mont(seed, ...)[elem_start:...].
    Raw limbs are generated on the GPU and reduced / converted in place by the library."""
    dev = torch.device("cuda", device_index)
    out = torch.empty((n_elems, 4), dtype=torch.int64, device=dev)
    for s in range(0, n_elems, chunk):
        e = min(n_elems, s + chunk)
        out[s:e] = splitmix64_torch(torch, seed, 4 * (elem_start + s), 4 * (e - s), dev).view(e - s, 4)
    st = torch.cuda.current_stream(dev).cuda_stream
    native.check(native.lib.cpb_field_to_montgomery_dev(field_id, device_index, out.data_ptr(), out.data_ptr(), n_elems, st))
    return out


def bytes_torch(torch, seed: int, byte_start: int, nbytes: int, device):
    assert byte_start % 8 == 0 and nbytes % 8 == 0
    return splitmix64_torch(torch, seed, byte_start // 8, nbytes // 8, device).view(torch.uint8)


def randomness_torch(torch, seed: int, start: int, n: int, device):
    r = bytes_torch(torch, seed, 32 * start, 32 * n, device).view(n, 32).clone()
    r[:, 31] &= 0x07
    return r


class StreamRng:
    """Sequential reader of stream(seed) with the `.field(q)` interface the Pedersen setup mirrors take
    (crypto_primitives_b200.crh.pedersen.CRH.setup): four outputs -> little-endian limbs -> mod q."""

    def __init__(self, seed: int):
        self.seed, self.pos = seed, 0

    def next(self) -> int:
        v = int(splitmix64_np(self.seed, self.pos, 1)[0])
        self.pos += 1
        return v

    def field(self, q: int) -> int:
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        return v % q
