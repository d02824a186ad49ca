"""The experimental FP64-pipe field arithmetic (tools/fp52.cuh, not used by the library -- it measured slower than the
IMAD path, see its header: 52-bit limbs in doubles, radix 2^260, lazily reduced values) compiled
for the CPU with fma_rz emulated exactly and its range assertions enabled, against Python integers: random, extreme
(all limbs 2^52-1, values up to 4p), patterned operands, and long mul/sqr/add chains that must stay within the lazy bounds."""
import ctypes as C
import random

import numpy as np
import pytest

from helpers import build_host_shim
from oracle.fields import MODULI

R = 1 << 260


@pytest.fixture(scope="module")
def shim():
    return build_host_shim("fp52_host_shim")


def _w(x):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]


def _v(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


def operands(rnd, p, n, top):
    """values < top*p: random, limb-saturated, near multiples of p."""
    lim = min(top * p, 1 << 256)
    out = []
    for i in range(n):
        k = i % 5
        if k == 0:
            v = rnd.randrange(lim)
        elif k == 1:
            v = sum(rnd.choice((0, (1 << 52) - 1, 1, (1 << 52) - 2, rnd.getrandbits(52))) << (52 * j) for j in range(5)) % lim
        elif k == 2:
            v = (rnd.randrange(1, top + 1) * p - 1 - rnd.getrandbits(rnd.choice((1, 20, 60)))) % lim
        elif k == 3:
            v = (rnd.randrange(0, top) * p + rnd.getrandbits(rnd.choice((1, 20, 60)))) % lim
        else:
            v = lim - 1 - rnd.getrandbits(40)
        out.append(v)
    return out


def run(shim, fid, which, A, B, Cc=None, chain=0):
    n = len(B) if which != 3 else len(A) // 3
    a = np.array([_w(x) for x in A], dtype=np.uint32)
    b = np.array([_w(x) for x in B], dtype=np.uint32)
    c = np.array([_w(x) for x in (Cc or [0] * n)], dtype=np.uint32)
    out = np.zeros((n, 8), dtype=np.uint32)
    shim.fp52_host_op(fid, which, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                      out.ctypes.data_as(C.c_void_p), C.c_long(n), C.c_int(chain))
    return [_v(x) for x in out]


@pytest.mark.parametrize("fid,name", list(enumerate(MODULI)))
def test_mul_sqr_add_dot(shim, fid, name):
    p = MODULI[name]
    rnd = random.Random(52 + fid)
    rinv = pow(R, -1, p)
    n = 4000
    A, B = operands(rnd, p, n, 4), operands(rnd, p, n, 4)
    rnd.shuffle(B)
    assert run(shim, fid, 0, A, B) == [x * y * rinv % p for x, y in zip(A, B)]
    assert run(shim, fid, 1, A, A) == [x * x * rinv % p for x in A]
    # additions: one operand a state value (< 2p + 2^208), the other a constant / product (< 1.34 p)
    S = [min(v, 2 * p + (1 << 208) - 1) for v in operands(rnd, p, n, 2)]
    Cn = [v % (p + p // 3) for v in operands(rnd, p, n, 2)]
    assert run(shim, fid, 2, S, Cn) == [(x + y) % p for x, y in zip(S, Cn)]
    # three-term dot product with an addend
    m = 1500
    A3, B3 = operands(rnd, p, 3 * m, 2), operands(rnd, p, 3 * m, 1)
    Cc = operands(rnd, p, m, 2)
    exp = [(sum(A3[3 * i + j] * B3[3 * i + j] for j in range(3)) * rinv + Cc[i]) % p for i in range(m)]
    assert run(shim, fid, 3, A3, B3, Cc) == exp


@pytest.mark.parametrize("fid,name", list(enumerate(MODULI)))
def test_lazy_chains_stay_in_range(shim, fid, name):
    """x <- (x*b or x^2) + b, 200 times without canonicalising: the range assertions of the emulation (limbs < 2^52,
    columns < 2^63, values < 4p) must hold throughout and the end result must be exact."""
    p = MODULI[name]
    rnd = random.Random(99 + fid)
    rinv = pow(R, -1, p)
    n, chain = 300, 200
    X, B = operands(rnd, p, n, 2), [v % (p + p // 3) for v in operands(rnd, p, n, 2)]
    exp = []
    for x, b in zip(X, B):
        for k in range(chain):
            x = ((x * x if k & 1 else x * b) * rinv + b) % p
        exp.append(x)
    assert run(shim, fid, 4, X, B, chain=chain) == exp
