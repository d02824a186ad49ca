// CPU build of the device Poseidon code path (crypto_primitives_b200/csrc/poseidon.cuh with the
// PTX primitives emulated) driven by the product's own host-side schedule derivation
// (poseidon_host.hpp).  Lets tests/test_poseidon_host.py check, without a GPU, that the sparse
// round schedule + device permutation reproduce the oracle bit-for-bit.  Not part of the product.
#include "../../crypto_primitives_b200/csrc/poseidon_team.cuh"
#include "../../crypto_primitives_b200/csrc/poseidon_host.hpp"
#include <cstring>
#include <vector>
using namespace cpb;

static PoseidonDev to_dev(const host::PoseidonSchedule& S) {
    PoseidonDev D;
    D.t = S.t; D.rate = S.rate; D.cap = S.capacity; D.rf = S.rf; D.rp = S.rp; D.sparse = S.sparse; D.alpha = S.alpha;
    D.off_c = S.off_c; D.off_m = S.off_m; D.off_mpre = S.off_mpre; D.off_cp0 = S.off_cp0; D.off_pc = S.off_pc;
    D.off_sp = S.off_sp; D.off_arkp = S.off_arkp; D.off_mod = S.off_mod; D.off_sc0 = S.off_sc0; D.n_elems = S.n_elems; D.zero = 0;
    return D;
}

template <class F, int T>
static void run_sponge(const PoseidonDev& D, const u32* cs, const u32* in, long len, long n_out, long n, u32* out) {
    u32 pm[8];
    ld_elem(pm, cs + 8 * D.off_mod);
    const bool single = len <= D.rate && n_out >= 1 && n_out <= D.rate && D.cap >= 1;      // as launch_crh_ft selects the kernel
    for (long i = 0; i < n; i++) {
        if (single) pos_hash_single<F, T>(out + 8 * n_out * i, (int)n_out, in + 8 * len * i, (int)len, D, cs, pm);
        else pos_sponge<F, T>(out + 8 * n_out * i, n_out, in + 8 * len * i, len, D, cs, pm);
    }
}
template <class F, int T>
static void run_verify(const PoseidonDev& D, const u32* cs, const u32* root, const u32* leaves, long leaf_len, const u32* sib,
                       const u32* paths, int plen, const unsigned long long* idx, unsigned char* ok, long n) {
    u32 pm[8];
    ld_elem(pm, cs + 8 * D.off_mod);
    for (long i = 0; i < n; i++)
        ok[i] = pos_verify_path<F, T>(leaves + 8 * leaf_len * i, leaf_len, sib + 8 * i, paths + 8 * (long)plen * i, plen, idx[i], root, D, cs, D, cs, pm);
}

template <class F, int T>
static void run(const PoseidonDev& D, const u32* cs, const u32* in, long len, long n, u32* out) {
    u32 pm[8];
    ld_elem(pm, cs + 8 * D.off_mod);
    const bool single = len <= D.rate && D.cap >= 1;                                       // as launch_crh_ft selects the kernel
    for (long i = 0; i < n; i++) {
        if (single) pos_hash_single<F, T>(out + 8 * i, 1, in + 8 * len * i, (int)len, D, cs, pm);
        else pos_crh<F, T>(out + 8 * i, in + 8 * len * i, len, D, cs, pm);
    }
}

template <class F>
static int run_t(const PoseidonDev& D, const u32* cs, const u32* in, long len, long n, u32* out) {
    switch (D.t) {
        case 2: run<F, 2>(D, cs, in, len, n, out); return 0;
        case 3: run<F, 3>(D, cs, in, len, n, out); return 0;
        case 4: run<F, 4>(D, cs, in, len, n, out); return 0;
        case 5: run<F, 5>(D, cs, in, len, n, out); return 0;
        case 9: run<F, 9>(D, cs, in, len, n, out); return 0;
    }
    return 1;
}

// returns: -1 error, else the `sparse` flag of the derived schedule
extern "C" int host_poseidon_crh(int field, int rate, int cap, int rf, int rp, unsigned long long alpha,
                                 const uint64_t* ark, const uint64_t* mds, int allow_sparse,
                                 const uint64_t* in, long len, long n, uint64_t* out) {
    const uint64_t* mod = host::field_modulus(field);
    if (!mod) return -1;
    host::Field F(mod);
    host::PoseidonParams P;
    P.rate = rate; P.capacity = cap; P.full_rounds = rf; P.partial_rounds = rp; P.alpha = alpha;
    int t = rate + cap;
    P.ark.resize((size_t)(rf + rp) * t);
    P.mds.resize((size_t)t * t);
    memcpy(P.ark.data(), ark, P.ark.size() * 32);
    memcpy(P.mds.data(), mds, P.mds.size() * 32);
    host::PoseidonSchedule S = host::derive_schedule(F, P, allow_sparse != 0);
    PoseidonDev D = to_dev(S);
    const u32* cs = reinterpret_cast<const u32*>(S.consts.data());
    const u32* i32 = reinterpret_cast<const u32*>(in);
    u32* o32 = reinterpret_cast<u32*>(out);
    int rc = 1;
    switch (field) {
        case 0: rc = run_t<Bls12_381_Fr>(D, cs, i32, len, n, o32); break;
        case 1: rc = run_t<Bn254_Fr>(D, cs, i32, len, n, o32); break;
        case 2: rc = run_t<Jubjub_Fr>(D, cs, i32, len, n, o32); break;
        case 3: rc = run_t<Bls12_377_Fr>(D, cs, i32, len, n, o32); break;
    }
    return rc ? -1 : S.sparse;
}

// find_poseidon_ark_and_mds on the product's host code; outputs Montgomery limbs.
extern "C" int host_poseidon_find_params(int field, int rate, int rf, int rp, int skip, uint64_t* ark, uint64_t* mds) {
    const uint64_t* mod = host::field_modulus(field);
    if (!mod) return -1;
    host::Field F(mod);
    host::FeVec a, m;
    host::find_poseidon_ark_and_mds(F, (uint64_t)F.bits, rate, rf, rp, skip, a, m);
    memcpy(ark, a.data(), a.size() * 32);
    memcpy(mds, m.data(), m.size() * 32);
    return 0;
}

static host::PoseidonSchedule make_schedule(int field, int rate, int cap, int rf, int rp, unsigned long long alpha,
                                            const uint64_t* ark, const uint64_t* mds) {
    host::Field F(host::field_modulus(field));
    host::PoseidonParams P;
    P.rate = rate; P.capacity = cap; P.full_rounds = rf; P.partial_rounds = rp; P.alpha = alpha;
    int t = rate + cap;
    P.ark.resize((size_t)(rf + rp) * t);
    P.mds.resize((size_t)t * t);
    memcpy(P.ark.data(), ark, P.ark.size() * 32);
    memcpy(P.mds.data(), mds, P.mds.size() * 32);
    return host::derive_schedule(F, P, true);
}

// absorb `len` then squeeze `n_out` native elements (t = 3 and 5, BLS12-381 Fr / BN254 Fr only: enough for the state machine)
extern "C" int host_poseidon_sponge(int field, int rate, int cap, int rf, int rp, unsigned long long alpha, const uint64_t* ark,
                                    const uint64_t* mds, const uint64_t* in, long len, long n_out, long n, uint64_t* out) {
    host::PoseidonSchedule S = make_schedule(field, rate, cap, rf, rp, alpha, ark, mds);
    PoseidonDev D = to_dev(S);
    const u32* cs = reinterpret_cast<const u32*>(S.consts.data());
    const u32* i32 = reinterpret_cast<const u32*>(in);
    u32* o32 = reinterpret_cast<u32*>(out);
    if (field == 0 && D.t == 3) run_sponge<Bls12_381_Fr, 3>(D, cs, i32, len, n_out, n, o32);
    else if (field == 0 && D.t == 5) run_sponge<Bls12_381_Fr, 5>(D, cs, i32, len, n_out, n, o32);
    else if (field == 1 && D.t == 3) run_sponge<Bn254_Fr, 3>(D, cs, i32, len, n_out, n, o32);
    else return -1;
    return 0;
}

extern "C" int host_poseidon_verify(int field, int rate, int cap, int rf, int rp, unsigned long long alpha, const uint64_t* ark,
                                    const uint64_t* mds, const uint64_t* root, const uint64_t* leaves, long leaf_len,
                                    const uint64_t* sib, const uint64_t* paths, int plen, const unsigned long long* idx,
                                    unsigned char* ok, long n) {
    host::PoseidonSchedule S = make_schedule(field, rate, cap, rf, rp, alpha, ark, mds);
    PoseidonDev D = to_dev(S);
    const u32* cs = reinterpret_cast<const u32*>(S.consts.data());
    if (field == 0 && D.t == 3)
        run_verify<Bls12_381_Fr, 3>(D, cs, (const u32*)root, (const u32*)leaves, leaf_len, (const u32*)sib, (const u32*)paths, plen, idx, ok, n);
    else if (field == 2 && D.t == 3)
        run_verify<Jubjub_Fr, 3>(D, cs, (const u32*)root, (const u32*)leaves, leaf_len, (const u32*)sib, (const u32*)paths, plen, idx, ok, n);
    else return -1;
    return 0;
}

// CPU model of the four-warp team kernel (poseidon_team.cuh): each phase runs for w = 0..3 in sequence, i.e. the barrier
// after every phase is modelled exactly; slots start poisoned so that a read of a never-written slot shows.
template <class F> static void run_team(const PoseidonDev& D, const u32* cs, const u32* pairs, long n, u32* out) {
    u32 pm[8];
    ld_elem(pm, cs + 8 * D.off_mod);
    int tb_alpha = 0, tb_e = 0;
    for (int i = 63; i > 0; i--) if ((D.alpha >> i) & 1) { tb_alpha = i; break; }
    for (int i = 63; i > 0; i--) if (((D.alpha - 1) >> i) & 1) { tb_e = i; break; }
    std::vector<u32> xb(kTeamXbWords, 0xdeadbeefu);
    for (long i = 0; i < n; i++) {
        u32 s[4][8], t[4][8];
        for (int w = 0; w < 4; w++) { fp_zero(s[w]); fp_zero(t[w]); }
        for (int j = 0; j < 8; j++) { s[1][j] = pairs[16 * i + j]; s[2][j] = pairs[16 * i + 8 + j]; }
        const int lane = (int)(i & 31);
        for (int r = 0; r < D.rf + D.rp; r++) {
            for (int w = 0; w < 4; w++) team_phase1<F>(s[w], t[w], w, lane, r, D, cs, pm, xb.data(), tb_alpha, tb_e);
            for (int w = 0; w < 4; w++) team_phase2<F>(s[w], t[w], w, lane, r, D, cs, pm, xb.data());
            for (int w = 0; w < 4; w++) team_publish<F>(s[w], w, lane, r, D, cs, xb.data());
        }
        for (int j = 0; j < 8; j++) out[8 * i + j] = s[1][j];
    }
}
extern "C" int host_poseidon_team_compress(int field, int rf, int rp, unsigned long long alpha, const uint64_t* ark, const uint64_t* mds,
                                           int allow_sparse, const uint64_t* pairs, long n, uint64_t* out) {
    host::Field F(host::field_modulus(field));
    host::PoseidonParams P;
    P.rate = 2; P.capacity = 1; P.full_rounds = rf; P.partial_rounds = rp; P.alpha = alpha;
    P.ark.resize((size_t)(rf + rp) * 3);
    P.mds.resize(9);
    memcpy(P.ark.data(), ark, P.ark.size() * 32);
    memcpy(P.mds.data(), mds, P.mds.size() * 32);
    host::PoseidonSchedule S = host::derive_schedule(F, P, allow_sparse != 0);
    PoseidonDev D = to_dev(S);
    const u32* cs = reinterpret_cast<const u32*>(S.consts.data());
    switch (field) {
        case 0: run_team<Bls12_381_Fr>(D, cs, (const u32*)pairs, n, (u32*)out); break;
        case 1: run_team<Bn254_Fr>(D, cs, (const u32*)pairs, n, (u32*)out); break;
        case 2: run_team<Jubjub_Fr>(D, cs, (const u32*)pairs, n, (u32*)out); break;
        case 3: run_team<Bls12_377_Fr>(D, cs, (const u32*)pairs, n, (u32*)out); break;
        default: return -1;
    }
    return S.sparse;
}
