// Device-vs-host-emulation bisect harness for the templated field/Poseidon code (development tool).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../crypto_primitives_b200/csrc/poseidon.cuh"
using namespace cpb;

template <class F, int T> __global__ void k_dot(const u32* a, const u32* b, u32* out, int n, int zero) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s[T][8], pm[8], d[8];
    fp_modulus<F>(pm);
    for (int j = 0; j < 8; j++) pm[j] += threadIdx.x * zero;
#pragma unroll
    for (int t = 0; t < T; t++) ld_elem(s[t], a + 8 * (T * i + t));
    fp_dot<F, T>(d, s, b + 8 * T * (i % 4) + threadIdx.x * zero, pm);
    st_elem(out + 8 * i, d);
}
template <class F, int T> __global__ void k_rot(const u32* a, u32* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s[T][8];
#pragma unroll
    for (int t = 0; t < T; t++) ld_elem(s[t], a + 8 * (T * i + t));
#pragma unroll 1
    for (int r = 0; r < (i % T) + 1; r++) pos_rotl<T>(s);
#pragma unroll
    for (int t = 0; t < T; t++) st_elem(out + 8 * (T * i + t), s[t]);
}
template <class F, int T> void run(const char* name) {
    const int n = 256;
    std::vector<u32> a(n * T * 8), b(4 * T * 8), out(n * 8), ref(n * 8), ro(n * T * 8), rr(n * T * 8);
    srand(T);
    auto rnd_elem = [](u32* e) { for (int j = 0; j < 8; j++) e[j] = (u32)rand() * 2654435761u + rand(); e[7] &= 0x3fffffffu; };
    for (int i = 0; i < n * T; i++) rnd_elem(&a[8 * i]);
    for (int i = 0; i < 4 * T; i++) rnd_elem(&b[8 * i]);
    for (int j = 0; j < 8; j++) { a[j] = F::P(j); b[j] = F::P(j); }
    a[0] -= 1; b[0] -= 1;   // p-1
    u32 *da, *db, *dout, *dro;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, out.size() * 4); cudaMalloc(&dro, ro.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    k_dot<F, T><<<(n + 63) / 64, 64>>>(da, db, dout, n, 0);
    k_rot<F, T><<<(n + 63) / 64, 64>>>(da, dro, n);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ro.data(), dro, ro.size() * 4, cudaMemcpyDeviceToHost);
    printf("%s T=%d cuda: %s\n", name, T, cudaGetErrorString(cudaGetLastError()));
    int bad = 0, badr = 0;
    for (int i = 0; i < n; i++) {
        u32 s[T][8], pm[8], d[8];
        fp_modulus<F>(pm);
        for (int t = 0; t < T; t++) for (int j = 0; j < 8; j++) s[t][j] = a[8 * (T * i + t) + j];
        fp_dot<F, T>(d, s, &b[8 * T * (i % 4)], pm);
        for (int j = 0; j < 8; j++) if (d[j] != out[8 * i + j]) { bad++; break; }
        for (int r = 0; r < (i % T) + 1; r++) pos_rotl<T>(s);
        for (int t = 0; t < T; t++) for (int j = 0; j < 8; j++) if (s[t][j] != ro[8 * (T * i + t) + j]) { badr++; t = T; break; }
    }
    printf("%s T=%d fp_dot mismatches %d / %d ; rotl mismatches %d\n", name, T, bad, n, badr);
}
int main() {
    run<Bls12_381_Fr, 3>("bls"); run<Bls12_381_Fr, 4>("bls"); run<Bls12_381_Fr, 5>("bls"); run<Bls12_381_Fr, 6>("bls"); run<Bls12_381_Fr, 9>("bls");
    run<Bn254_Fr, 4>("bn254"); run<Bn254_Fr, 5>("bn254");
    return 0;
}
