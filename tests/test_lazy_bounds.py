"""The inequalities behind the lazy reduction of the BN254 Fr alpha = 5 rounds (csrc/poseidon.cuh: LZ / F::LAZY5, csrc/fp.cuh:
fp_mul<F,LAZY>, fp_dot<F,T,EX>), checked with exact rationals for the real modulus.  A Montgomery product of a < A*p and
b < B*p returns (a*b + M*p)/R < p*(A*B*p/R + 1); the accumulators of a product need (full operand) + p < R, those of a
T-term lazy dot without its overflow word need sum_j a_j + p <= R."""
from fractions import Fraction as Fr

P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001      # BN254 Fr
R = 1 << 256
rho = Fr(P, R)


def mont(a, b):
    """upper bound (in units of p) of the unreduced Montgomery product of values below a*p and b*p"""
    return a * b * rho + 1


def test_field_flag_condition():
    top = P >> 224
    assert 100 * (top + 1) <= 19 * (1 << 32)                 # the static_assert of poseidon.cuh: p/R <= 0.19
    assert rho < Fr(19, 100)
    assert 5 * (top + 1) <= (1 << 32) < 6 * (top + 1)        # dot_needs_x<F, 4> false, <F, 5> true: T <= 3 for LZ


def test_partial_round_chain():
    x = Fr(2)                                                # x = d + c, d and c canonical, not reduced
    assert x * rho < 1                                       # fits 256 bits
    x2 = mont(x, x)
    x4 = mont(x2, x2)
    assert x4 + 1 < 1 / rho                                  # fp_mul(s0, x4, x): full operand x4, x4 + p < R
    y = mont(x4, x)
    assert max(x2, x4, y) < 2                                # every lazy value stays below 2p (closed: 4p/R + 1 <= 2)
    for t in (2, 3):
        assert (y + (t - 1) + 1) * rho <= 1                  # row product: y + (t-1) canonical lanes + p <= R, no overflow word
        assert y + (t - 1) < t + 1                           # ... and within what EX = 1 declares
        assert (y + (t - 1)) * rho + 1 <= 2                  # its result < 2p: one conditional subtraction gives the canonical d
    assert y + 1 < 1 / rho                                   # column products v_j * y: full operand y, y + p < R
    assert mont(y, 1) < 2                                    # and they are reduced to canonical by the ordinary final subtraction


def test_full_round_chain():
    x2 = mont(1, 1)                                          # canonical lane after the constant addition
    x4 = mont(x2, x2)
    assert x4 + 1 < 1 / rho
    y = mont(x4, 1)
    assert x2 < Fr(119, 100) and x4 < Fr(127, 100) and y < Fr(124, 100)
    for t in (2, 3, 4):
        assert t * y < t + 1                                 # dense rows: t unreduced lanes count as t + 1 canonical terms (EX = 1)
        assert (t * y) * rho + 1 <= 2                        # one conditional subtraction suffices
    for t in (2, 3):
        assert (t * y + 1) * rho <= 1                        # no overflow word for the widths LZ admits
