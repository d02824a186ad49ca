"""Poseidon sponge / CRH oracle (Python big-int; test infrastructure only).

Restates, line by line:
  R/sponge/poseidon/grain_lfsr.rs:16-181   PoseidonGrainLFSR
  R/sponge/poseidon/traits.rs:69-146       default parameters, find_poseidon_ark_and_mds
  R/sponge/poseidon/mod.rs:66-345          PoseidonSponge (permute, absorb, squeeze)
  R/sponge/absorb.rs:108-122,154-167,284-292,334-342   Absorb for Fp / &[Fp] (identity)
  R/crh/poseidon/mod.rs:15-80              CRH, TwoToOneCRH
(R = /root/reference/crypto-primitives/src).  Elements are plain ints in [0,p).
"""
from __future__ import annotations

from dataclasses import dataclass, field as _field

from .fields import inv


class PoseidonGrainLFSR:
    """grain_lfsr.rs:16-181."""

    def __init__(self, is_sbox_an_inverse: bool, prime_num_bits: int, state_len: int,
                 num_full_rounds: int, num_partial_rounds: int):
        st = [False] * 80
        st[1] = True                                   # :25  b0,b1 describe the field
        st[5] = bool(is_sbox_an_inverse)               # :28-32

        def put(lo, hi, val):                          # :35-68  MSB first into b[lo..hi]
            cur = val
            for i in range(hi, lo - 1, -1):
                st[i] = (cur & 1) == 1
                cur >>= 1
        put(6, 17, prime_num_bits)
        put(18, 29, state_len)
        put(30, 39, num_full_rounds)
        put(40, 49, num_partial_rounds)
        for i in range(50, 80):                        # :71-73
            st[i] = True
        self.prime_num_bits = prime_num_bits
        self.state = st
        self.head = 0
        for _ in range(160):                           # init :177-181
            self.update()

    def update(self) -> bool:                          # :162-175
        s, h = self.state, self.head
        nb = (s[(h + 62) % 80] ^ s[(h + 51) % 80] ^ s[(h + 38) % 80]
              ^ s[(h + 23) % 80] ^ s[(h + 13) % 80] ^ s[h])
        s[h] = nb
        self.head = (h + 1) % 80
        return nb

    def get_bits(self, num_bits: int) -> list[bool]:   # :87-107
        res = []
        for _ in range(num_bits):
            new_bit = self.update()
            while not new_bit:
                self.update()
                new_bit = self.update()
            res.append(self.update())
        return res

    def _int_msb_first(self) -> int:
        v = 0
        for b in self.get_bits(self.prime_num_bits):   # first bit = most significant
            v = (v << 1) | int(b)
        return v

    def get_field_elements_rejection_sampling(self, num_elems: int, p: int) -> list[int]:  # :109-134
        assert p.bit_length() == self.prime_num_bits
        res = []
        for _ in range(num_elems):
            while True:
                v = self._int_msb_first()
                if v < p:                              # F::from_bigint returns None when >= p
                    res.append(v)
                    break
        return res

    def get_field_elements_mod_p(self, num_elems: int, p: int) -> list[int]:  # :136-160
        assert p.bit_length() == self.prime_num_bits
        return [self._int_msb_first() % p for _ in range(num_elems)]


def find_poseidon_ark_and_mds(p: int, prime_bits: int, rate: int, full_rounds: int,
                              partial_rounds: int, skip_matrices: int):
    """traits.rs:105-146."""
    t = rate + 1
    lfsr = PoseidonGrainLFSR(False, prime_bits, t, full_rounds, partial_rounds)
    ark = [lfsr.get_field_elements_rejection_sampling(t, p)
           for _ in range(full_rounds + partial_rounds)]
    for _ in range(skip_matrices):
        lfsr.get_field_elements_mod_p(2 * t, p)
    xs = lfsr.get_field_elements_mod_p(t, p)
    ys = lfsr.get_field_elements_mod_p(t, p)
    mds = [[inv((xs[i] + ys[j]) % p, p) for j in range(t)] for i in range(t)]
    return ark, mds


@dataclass
class PoseidonConfig:
    """sponge/poseidon/mod.rs:26-45 (+ the modulus, which the reference carries in the type)."""
    p: int
    full_rounds: int
    partial_rounds: int
    alpha: int
    ark: list
    mds: list
    rate: int
    capacity: int

    def __post_init__(self):                           # PoseidonConfig::new asserts :189-217
        t = self.rate + self.capacity
        assert len(self.ark) == self.full_rounds + self.partial_rounds
        assert all(len(r) == t for r in self.ark)
        assert len(self.mds) == t and all(len(r) == t for r in self.mds)


# R/sponge/test.rs:13-32 -- (rate, alpha, full, partial, skip) for the BLS12-381 Fr test field.
PARAMS_OPT_FOR_CONSTRAINTS = [(2, 17, 8, 31, 0), (3, 5, 8, 56, 0), (4, 5, 8, 56, 0), (5, 5, 8, 57, 0),
                              (6, 5, 8, 57, 0), (7, 5, 8, 57, 0), (8, 5, 8, 57, 0)]
PARAMS_OPT_FOR_WEIGHTS = [(r, 257, 8, 13, 0) for r in range(2, 9)]


def get_default_poseidon_parameters(p: int, rate: int, optimized_for_weights: bool,
                                    table=None) -> PoseidonConfig | None:
    """traits.rs:69-103 with the entry tables of sponge/test.rs (BLS12-381 Fr)."""
    if table is None:
        table = PARAMS_OPT_FOR_WEIGHTS if optimized_for_weights else PARAMS_OPT_FOR_CONSTRAINTS
    for (r, alpha, rf, rp, skip) in table:
        if r == rate:
            ark, mds = find_poseidon_ark_and_mds(p, p.bit_length(), rate, rf, rp, skip)
            return PoseidonConfig(p, rf, rp, alpha, ark, mds, rate, 1)
    return None


def permute(cfg: PoseidonConfig, state: list[int]) -> list[int]:
    """mod.rs:66-121: ark -> s-box (all lanes / lane 0) -> dense MDS, every round."""
    p, t = cfg.p, cfg.rate + cfg.capacity
    half = cfg.full_rounds // 2
    st = list(state)
    for r in range(cfg.full_rounds + cfg.partial_rounds):
        st = [(st[i] + cfg.ark[r][i]) % p for i in range(t)]                # apply_ark :79-83
        if r < half or r >= half + cfg.partial_rounds:                        # apply_s_box :66-77
            st = [pow(x, cfg.alpha, p) for x in st]
        else:
            st[0] = pow(st[0], cfg.alpha, p)
        st = [sum(st[j] * cfg.mds[i][j] for j in range(t)) % p for i in range(t)]  # apply_mds :85-96
    return st


class PoseidonSponge:
    """mod.rs:47-63, 124-186, 220-345.  mode = ('A', next_absorb_index) | ('S', next_squeeze_index)."""

    def __init__(self, cfg: PoseidonConfig):
        self.cfg = cfg
        self.state = [0] * (cfg.rate + cfg.capacity)
        self.mode = ("A", 0)

    def _permute(self):
        self.state = permute(self.cfg, self.state)

    def _absorb_internal(self, rate_start: int, elems: list[int]):           # :124-153
        c = self.cfg
        rem = list(elems)
        while True:
            if rate_start + len(rem) <= c.rate:
                for i, e in enumerate(rem):
                    k = c.capacity + i + rate_start
                    self.state[k] = (self.state[k] + e) % c.p
                self.mode = ("A", rate_start + len(rem))
                return
            n = c.rate - rate_start
            for i, e in enumerate(rem[:n]):
                k = c.capacity + i + rate_start
                self.state[k] = (self.state[k] + e) % c.p
            self._permute()
            rem = rem[n:]
            rate_start = 0

    def _squeeze_internal(self, rate_start: int, n_out: int) -> list[int]:   # :156-186
        c = self.cfg
        out = []
        rem = n_out
        while True:
            if rate_start + rem <= c.rate:
                out += self.state[c.capacity + rate_start: c.capacity + rate_start + rem]
                self.mode = ("S", rate_start + rem)
                return out
            n = c.rate - rate_start
            out += self.state[c.capacity + rate_start: c.capacity + rate_start + n]
            rem -= n
            if rem != 0:
                self._permute()
            rate_start = 0

    def absorb(self, elems: list[int]):                                      # :236-257
        elems = [e % self.cfg.p for e in elems]       # field_cast is the identity for native elements
        if not elems:
            return
        kind, idx = self.mode
        if kind == "A":
            if idx == self.cfg.rate:
                self._permute()
                idx = 0
            self._absorb_internal(idx, elems)
        else:
            self._absorb_internal(0, elems)

    def squeeze_native_field_elements(self, n: int) -> list[int]:            # :323-345
        kind, idx = self.mode
        if kind == "A":
            self._permute()
            return self._squeeze_internal(0, n)
        if idx == self.cfg.rate:
            self._permute()
            idx = 0
        return self._squeeze_internal(idx, n)


def crh_evaluate(cfg: PoseidonConfig, inp: list[int]) -> int:
    """crh/poseidon/mod.rs:30-40."""
    s = PoseidonSponge(cfg)
    s.absorb(list(inp))
    return s.squeeze_native_field_elements(1)[0]


def two_to_one_compress(cfg: PoseidonConfig, left: int, right: int) -> int:
    """crh/poseidon/mod.rs:66-79 (evaluate :58-64 is an alias)."""
    s = PoseidonSponge(cfg)
    s.absorb([left])
    s.absorb([right])
    return s.squeeze_native_field_elements(1)[0]
