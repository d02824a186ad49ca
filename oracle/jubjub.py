"""Twisted-Edwards curve oracle (Jubjub = ark_ed_on_bls12_381), affine, Python ints.

The reference's group arithmetic lives in ark-ec ^0.4 (`twisted_edwards::{Affine,
Projective}`, not under /root/reference); its call sites on the hot path are
R/crh/pedersen/mod.rs:50-53,116-124,128 and R/commitment/pedersen/mod.rs:88-104.
A group element has a unique affine representative, so the affine addition law
below is an exact restatement of `+=` followed by `into_affine()`.

Curve: a*x^2 + y^2 = 1 + d*x^2*y^2 over Fq, q = BLS12-381 Fr, a = -1,
d = -(10240/10241).  a is a square and d a non-square in Fq, so the addition law
is complete (no exceptional points) -- checked in tests/test_oracle_pedersen.py.
"""
from __future__ import annotations

from .fields import BLS12_381_FR, JUBJUB_FR, inv

Q = BLS12_381_FR
A = Q - 1
D = (-(10240 * inv(10241, Q))) % Q
ORDER = JUBJUB_FR          # prime-order subgroup
COFACTOR = 8
IDENTITY = (0, 1)


def is_on_curve(P) -> bool:
    x, y = P
    return (A * x * x + y * y - 1 - D * x * x % Q * y * y) % Q == 0


def add(P, R):
    x1, y1 = P
    x2, y2 = R
    k = D * x1 % Q * x2 % Q * y1 % Q * y2 % Q
    x3 = (x1 * y2 + y1 * x2) % Q * inv((1 + k) % Q, Q) % Q
    y3 = (y1 * y2 - A * x1 * x2) % Q * inv((1 - k) % Q, Q) % Q
    return (x3, y3)


def double(P):
    return add(P, P)


def neg(P):
    return ((-P[0]) % Q, P[1])


def mul(k: int, P):
    acc = IDENTITY
    base = P
    while k:
        if k & 1:
            acc = add(acc, base)
        base = double(base)
        k >>= 1
    return acc


def sqrt(n: int):
    """Tonelli-Shanks in Fq (q-1 = 2^32 * odd). Returns None for non-residues."""
    n %= Q
    if n == 0:
        return 0
    if pow(n, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 0, Q - 1
    while t % 2 == 0:
        s += 1
        t //= 2
    z = 2
    while pow(z, (Q - 1) // 2, Q) == 1:
        z += 1
    m, c, tt, r = s, pow(z, t, Q), pow(n, t, Q), pow(n, (t + 1) // 2, Q)
    while tt != 1:
        i, x = 0, tt
        while x != 1:
            x = x * x % Q
            i += 1
        b = pow(c, 1 << (m - i - 1), Q)
        m, c = i, b * b % Q
        tt, r = tt * c % Q, r * b % Q
    return r


def point_from_y(y: int):
    """x^2 = (1 - y^2) / (a - d*y^2); returns the root with even canonical x, or None."""
    y %= Q
    den = (A - D * y * y) % Q
    if den == 0:
        return None
    x = sqrt((1 - y * y) % Q * inv(den, Q) % Q)
    if x is None:
        return None
    if x & 1:
        x = Q - x
    return (x, y)


def serialize_uncompressed(P) -> bytes:
    """ark-serialize `serialize_uncompressed` of a TE affine point: x || y, 32-byte LE
    canonical each (dep, from memory; used by R/macros.rs:3-13, R/merkle_tree/mod.rs:71-78)."""
    return P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little")
