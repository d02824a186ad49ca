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


# ---- adversarial carry patterns -------------------------------------------------------------------------------
# Random operands exercise a carry out of an all-ones limb with probability ~2^-32 per limb, so a dropped carry in a
# rarely taken position survives any amount of random testing (round 1 shipped one in fp_sqr: the overflow word of a
# reduction row met T[i+8] = 0xffffffff and the carry was lost -- about one wrong squaring in 3*10^9, i.e. a wrong
# node in most 2^24-leaf trees).  These tests construct operands whose *products* and *intermediate sums* contain
# runs of 0xffffffff / 0 limbs at chosen positions.
PAT = [0x00000000, 0xFFFFFFFF, 0x00000001, 0xFFFFFFFE, 0x80000000, 0x7FFFFFFF]


def _pattern_value(rnd, nlimbs, density=0.6):
    v = 0
    for i in range(nlimbs):
        limb = rnd.choice(PAT) if rnd.random() < density else rnd.getrandbits(32)
        v |= limb << (32 * i)
    return v


def _crafted_square_roots(rnd, p, count):
    """a < p such that a*a has a run of all-ones limbs in its upper half (limbs 8..14)."""
    import math
    out = []
    top_bits = (p * p).bit_length()
    while len(out) < count:
        run = rnd.choice((1, 1, 2, 3))
        k = rnd.randrange(8, 15 - run + 1)
        if 32 * (k + run) >= top_bits - 2:
            continue
        v = rnd.randrange(1, (p * p) >> (32 * (k + run))) << (32 * (k + run))
        for j in range(run):
            v |= 0xFFFFFFFF << (32 * (k + j))
        v |= (0x80000000 | rnd.getrandbits(31)) << (32 * (k - 1))        # absorbs the isqrt remainder: no borrow into limb k
        v |= _pattern_value(rnd, k - 1, 0.3)
        a = math.isqrt(v)
        if a < p:
            out.append(a)
    return out


def _crafted_factor_pairs(rnd, p, count):
    """(a, b), both < p, with a*b = V - (V mod a) for a patterned V: all-ones / zero runs in the upper half; and pairs with
    a*b = V mod 2^256 exactly (patterned lower half, which drives the reduction multipliers m_i)."""
    out = []
    while len(out) < count:
        if rnd.random() < 0.5:
            a = rnd.randrange(1 << 200, p)
            v = _pattern_value(rnd, 16, 0.7) % (a * (p - 1))
            v |= 0x80000000 << (32 * 7)
            b = v // a
        else:
            a = rnd.randrange(1 << 200, p) | 1
            v = _pattern_value(rnd, 8, 0.8)
            b = v * pow(a, -1, 1 << 256) % (1 << 256)
        if 0 < b < p:
            out.append((a, b))
    return out


@pytest.mark.parametrize("fid,name", list(enumerate(MODULI)))
def test_adversarial_carry_patterns(shim, fid, name):
    p = MODULI[name]
    rnd = random.Random(100 + fid)
    rinv = pow(1 << 256, -1, p)

    def run(which, A, B):
        a = np.array([_limbs(x) for x in A], dtype=np.uint32)
        b = np.array([_limbs(x) for x in B], dtype=np.uint32)
        r = np.zeros_like(a)
        shim.fp_host_op(fid, which, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p),
                        C.c_ulonglong(0), C.c_long(len(A)))
        return [_val(x) for x in r]

    # squarings whose 512-bit square has all-ones limbs in the half that enters the reduction window limb by limb
    A = _crafted_square_roots(rnd, p, 6000)
    assert run(3, A, A) == [x * x * rinv % p for x in A]
    assert run(0, A, A) == [x * x * rinv % p for x in A]
    # products with patterned upper / lower halves
    pairs = _crafted_factor_pairs(rnd, p, 6000)
    A, B = [x for x, _ in pairs], [y for _, y in pairs]
    assert run(0, A, B) == [x * y * rinv % p for x, y in pairs]
    assert run(0, B, A) == [x * y * rinv % p for x, y in pairs]
    # operands that are themselves limb patterns (top limb kept below the modulus')
    P = []
    while len(P) < 8000:
        v = _pattern_value(rnd, 8, 0.85)
        if rnd.random() < 0.5:
            v = (v & ((1 << 224) - 1)) | (rnd.choice((p >> 224, (p >> 224) - 1, 0, 1)) << 224)
        if v < p:
            P.append(v)
    A, B = P[:4000], P[4000:]
    assert run(0, A, B) == [x * y * rinv % p for x, y in zip(A, B)]
    assert run(3, P, P) == [x * x * rinv % p for x in P]
    assert run(1, A, B) == [(x + y) % p for x, y in zip(A, B)]
    assert run(2, A, B) == [(x - y) % p for x, y in zip(A, B)]
    # near-modulus operands: results just below / above p before the final subtraction
    near = [p - 1 - rnd.getrandbits(rnd.choice((1, 8, 32, 64))) for _ in range(2000)]
    assert run(0, near, near[::-1]) == [x * y * rinv % p for x, y in zip(near, near[::-1])]
    assert run(3, near, near) == [x * x * rinv % p for x in near]


@pytest.mark.parametrize("fid,name", list(enumerate(MODULI)))
@pytest.mark.parametrize("t", [2, 3, 5, 9])
def test_lazy_dot_product_adversarial(shim, fid, name, t):
    """fp_dot<T> (one reduction for T products): random, patterned and maximal operands -- the running sum reaches
    (T+1) p 2^32 and spills into the overflow word."""
    p = MODULI[name]
    rnd = random.Random(200 + 10 * fid + t)
    rinv = pow(1 << 256, -1, p)
    n = 3000
    rows = []
    for i in range(n):
        kind = i % 4
        if kind == 0:
            a = [rnd.randrange(p) for _ in range(t)]; b = [rnd.randrange(p) for _ in range(t)]
        elif kind == 1:
            a = [p - 1 - rnd.getrandbits(8) for _ in range(t)]; b = [p - 1 - rnd.getrandbits(8) for _ in range(t)]
        elif kind == 2:
            a = [_pattern_value(rnd, 8, 0.85) % p for _ in range(t)]; b = [_pattern_value(rnd, 8, 0.85) % p for _ in range(t)]
        else:   # first product crafted, the rest patterned: sum with all-ones limbs
            x, y = _crafted_factor_pairs(rnd, p, 1)[0]
            a = [x] + [_pattern_value(rnd, 8, 0.9) % p for _ in range(t - 1)]
            b = [y] + [rnd.choice((0, 1, p - 1)) for _ in range(t - 1)]
        rows.append((a, b))
    A = np.array([[_limbs(x) for x in a] for a, _ in rows], dtype=np.uint32)
    B = np.array([[_limbs(x) for x in b] for _, b in rows], dtype=np.uint32)
    R = np.zeros((n, 8), dtype=np.uint32)
    shim.fp_host_dot(fid, t, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p), C.c_long(n))
    assert [_val(x) for x in R] == [sum(x * y for x, y in zip(a, b)) * rinv % p for a, b in rows]
