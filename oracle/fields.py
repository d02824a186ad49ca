"""Prime fields on the hot path (oracle; test infrastructure only).

The reference takes its fields from third-party crates that are NOT under
/root/reference (ark-ff ^0.4 `Fp<MontBackend<_,4>,4>`, ark-bls12-381,
ark-ed-on-bls12-381, ...; crypto-primitives/Cargo.toml:18-38,55-61).  Field
arithmetic has canonical results (unique representative < p), so plain Python
integers mod p are an exact restatement.

Interchange layout at the C-ABI: 4 x u64 little-endian limbs in
Montgomery form, R = 2^256, fully reduced -- what ark-ff's `Fp.0.0` holds.
"""
from __future__ import annotations

# BLS12-381 scalar field; the reference's own test field (R/sponge/test.rs:5-12).
BLS12_381_FR = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
# BN254 scalar field (BASELINE.json config 4; not in the reference tree).
BN254_FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
# Jubjub (ark_ed_on_bls12_381) scalar field; the Merkle field test hashes over it
# (R/merkle_tree/tests/mod.rs:195, test_utils.rs:4).
JUBJUB_FR = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7
# BLS12-377 scalar field (the CRH gadget test field, R/crh/poseidon/constraints.rs:133).
BLS12_377_FR = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001

FIELD_IDS = {"bls12_381_fr": 0, "bn254_fr": 1, "jubjub_fr": 2, "bls12_377_fr": 3}
MODULI = {
    "bls12_381_fr": BLS12_381_FR,
    "bn254_fr": BN254_FR,
    "jubjub_fr": JUBJUB_FR,
    "bls12_377_fr": BLS12_377_FR,
}

R_BITS = 256
MASK64 = (1 << 64) - 1


def modulus_bits(p: int) -> int:
    return p.bit_length()


def inv(a: int, p: int) -> int:
    return pow(a, p - 2, p)


def to_mont(x: int, p: int) -> int:
    return (x << R_BITS) % p


def from_mont(x: int, p: int) -> int:
    return (x * pow(1 << R_BITS, -1, p)) % p


def to_limbs(x: int) -> list[int]:
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def from_limbs(l) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


class SplitMix64:
    """Synthetic-input generator shared by oracle, tests and bench."""

    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def field(self, p: int) -> int:
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        return v % p

    def bytes(self, n: int) -> bytes:
        out = bytearray()
        while len(out) < n:
            out += self.next().to_bytes(8, "little")
        return bytes(out[:n])
