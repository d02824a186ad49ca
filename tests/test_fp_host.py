"""The device field arithmetic (csrc/fp.cuh) compiled for the CPU with the PTX carry-chain
primitives emulated (csrc/ptx.cuh), checked against Python integers.  Catches arithmetic bugs
before any GPU time is spent; the GPU tests then only have to prove the PTX path agrees."""
import ctypes as C
import random

import numpy as np
import pytest

from helpers import build_host_shim
from oracle.fields import MODULI


@pytest.fixture(scope="module")
def shim():
    return build_host_shim("fp_host_shim")


def _limbs(x):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def _val(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


@pytest.mark.parametrize("fid,name", list(enumerate(MODULI)))
def test_field_ops(shim, fid, name):
    p = MODULI[name]
    rnd = random.Random(fid)
    n = 800
    A = [rnd.randrange(p) for _ in range(n)]
    B = [rnd.randrange(p) for _ in range(n)]
    edge = [0, 1, 2, p - 1, p - 2, (1 << 255) % p, (1 << 256) % p, p >> 1,
            (0xDEADBEEF << 32) % p, (0x1234567 << 64) % p, (7 << 224) % p, (p - 1) & ~0xFFFFFFFF]   # zero low limbs: m == 0 rows
    for i, (x, y) in enumerate((x, y) for x in edge for y in edge):
        A[i], B[i] = x, y
    a = np.array([_limbs(x) for x in A], dtype=np.uint32)
    b = np.array([_limbs(x) for x in B], dtype=np.uint32)
    r = np.zeros_like(a)
    rinv = pow(1 << 256, -1, p)

    def run(which, alpha=0, nn=n):
        shim.fp_host_op(fid, which, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                        r.ctypes.data_as(C.c_void_p), C.c_ulonglong(alpha), C.c_long(nn))
        return [_val(x) for x in r[:nn]]

    assert run(0) == [x * y * rinv % p for x, y in zip(A, B)]
    assert run(1) == [(x + y) % p for x, y in zip(A, B)]
    assert run(2) == [(x - y) % p for x, y in zip(A, B)]
    assert run(3) == [x * x * rinv % p for x in A]
    r2 = pow(1 << 256, 2, p)
    assert run(4, 0, 40) == [pow(x, -1, p) * r2 % p if x else 0 for x in A[:40]]
    for alpha in (1, 2, 3, 5, 17, 31, 257):
        assert run(5, alpha, 100) == [(pow(x * rinv % p, alpha, p) << 256) % p for x in A[:100]]
