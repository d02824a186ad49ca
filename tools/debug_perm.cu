// Device-vs-host-emulation check of pos_permute<F,T> with a real derived schedule (development tool).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../crypto_primitives_b200/csrc/poseidon.cuh"
#include "../crypto_primitives_b200/csrc/poseidon_host.hpp"
using namespace cpb;

template <class F, int T> __global__ void k_perm(PoseidonDev P, const u32* consts, const u32* in, u32* out, int n, int stage) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32* ct = consts + threadIdx.x * P.zero;
    u32 pm[8], s[T][8];
    ld_elem(pm, ct + 8 * P.off_mod);
#pragma unroll
    for (int t = 0; t < T; t++) ld_elem(s[t], in + 8 * (T * i + t));
    if (stage == 0) pos_permute<F, T>(s, P, ct, pm);
    if (stage == 1) pos_add_vec<F, T>(s, ct + 8 * P.off_c);
    if (stage == 2) { for (int j = 0; j < T; j++) { pos_sbox<F>(s[0], P.alpha, 2, pm); pos_rotl<T>(s); } }
#pragma unroll
    for (int t = 0; t < T; t++) st_elem(out + 8 * (T * i + t), s[t]);
}
template <class F, int T> void run(int field_id, int rf, int rp) {
    host::Field HF(host::field_modulus(field_id));
    host::PoseidonParams PP;
    PP.rate = T - 1; PP.capacity = 1; PP.full_rounds = rf; PP.partial_rounds = rp; PP.alpha = 5;
    srand(T * 100 + rp);
    auto rnd = [&]() { uint64_t v[4]; for (int j = 0; j < 4; j++) v[j] = ((uint64_t)rand() << 33) ^ ((uint64_t)rand() << 11) ^ rand(); v[3] &= 0x0fffffffffffffffull; return HF.from_canonical(v); };
    for (int i = 0; i < (rf + rp) * T; i++) PP.ark.push_back(rnd());
    for (int i = 0; i < T * T; i++) PP.mds.push_back(rnd());
    host::PoseidonSchedule S = host::derive_schedule(HF, PP, true);
    PoseidonDev D;
    D.t = S.t; D.rate = S.rate; D.cap = S.capacity; D.rf = S.rf; D.rp = S.rp; D.sparse = S.sparse; D.alpha = S.alpha;
    D.off_c = S.off_c; D.off_m = S.off_m; D.off_mpre = S.off_mpre; D.off_cp0 = S.off_cp0; D.off_pc = S.off_pc;
    D.off_sp = S.off_sp; D.off_arkp = S.off_arkp; D.off_mod = S.off_mod; D.n_elems = S.n_elems; D.zero = 0;
    const int n = 128;
    std::vector<u32> in(n * T * 8), out(n * T * 8);
    for (int i = 0; i < n * T; i++) { host::Fe e = rnd(); memcpy(&in[8 * i], e.l, 32); }
    u32 *dc, *din, *dout;
    cudaMalloc(&dc, S.consts.size() * 8); cudaMalloc(&din, in.size() * 4); cudaMalloc(&dout, out.size() * 4);
    cudaMemcpy(dc, S.consts.data(), S.consts.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(din, in.data(), in.size() * 4, cudaMemcpyHostToDevice);
    const u32* cs = (const u32*)S.consts.data();
    for (int stage = 0; stage < 3; stage++) {
        k_perm<F, T><<<(n + 63) / 64, 64>>>(D, dc, din, dout, n, stage);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < n; i++) {
            u32 s[T][8], pm[8];
            ld_elem(pm, cs + 8 * D.off_mod);
            for (int t = 0; t < T; t++) for (int j = 0; j < 8; j++) s[t][j] = in[8 * (T * i + t) + j];
            if (stage == 0) pos_permute<F, T>(s, D, cs, pm);
            if (stage == 1) pos_add_vec<F, T>(s, cs + 8 * D.off_c);
            if (stage == 2) { for (int j = 0; j < T; j++) { pos_sbox<F>(s[0], D.alpha, 2, pm); pos_rotl<T>(s); } }
            bool ok = true;
            for (int t = 0; t < T; t++) for (int j = 0; j < 8; j++) if (s[t][j] != out[8 * (T * i + t) + j]) ok = false;
            bad += !ok;
        }
        printf("field %d T=%d rf=%d rp=%d sparse=%d stage %d (%s): mismatches %d / %d  [%s]\n", field_id, T, rf, rp, S.sparse, stage,
               stage == 0 ? "permute" : stage == 1 ? "add_vec" : "sbox+rotl", bad, n, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    run<Bls12_381_Fr, 4>(0, 2, 0); run<Bls12_381_Fr, 5>(0, 2, 0); run<Bls12_381_Fr, 5>(0, 2, 2); run<Bls12_381_Fr, 6>(0, 2, 0);
    run<Bn254_Fr, 5>(1, 2, 0);
    return 0;
}
