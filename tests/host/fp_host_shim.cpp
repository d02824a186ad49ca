// CPU build of crypto_primitives_b200/csrc/fp.cuh (PTX primitives emulated, see ptx.cuh) so the
// device field arithmetic can be checked bit-for-bit against Python integers without a GPU.
// Built and driven by tests/test_fp_host.py.  Not part of the product library.
#include "../../crypto_primitives_b200/csrc/fp.cuh"
#include <cstring>
using namespace cpb;

template <class F> static void op(int which, const u32* a, const u32* b, u32* r, unsigned long long alpha) {
    u32 x[8], y[8], z[8], pm[8];
    fp_modulus<F>(pm);
    memcpy(x, a, 32); memcpy(y, b, 32);
    switch (which) {
        case 0: fp_mul<F>(z, x, y, pm); break;
        case 1: fp_add<F>(z, x, y); break;
        case 2: fp_sub<F>(z, x, y); break;
        case 3: fp_sqr<F>(z, x, pm); break;
        case 4: fp_inv<F>(z, x, pm); break;
        case 5: fp_pow_alpha<F>(x, alpha, pm); memcpy(z, x, 32); break;
        default: memset(z, 0, 32);
    }
    memcpy(r, z, 32);
}

// r[i] = sum_j a[i][j] * b[i][j] / R  (T terms, operands interleaved per item)
template <class F, int T> static void dot(const u32* a, const u32* b, u32* r, long n) {
    u32 pm[8];
    fp_modulus<F>(pm);
    for (long i = 0; i < n; i++) {
        u32 x[T][8], z[8];
        memcpy(x, a + 8 * T * i, 32 * T);
        fp_dot<F, T>(z, x, b + 8 * T * i, pm);
        memcpy(r + 8 * i, z, 32);
    }
}
template <class F> static void dot_t(int t, const u32* a, const u32* b, u32* r, long n) {
    switch (t) {
        case 2: dot<F, 2>(a, b, r, n); break;
        case 3: dot<F, 3>(a, b, r, n); break;
        case 5: dot<F, 5>(a, b, r, n); break;
        case 9: dot<F, 9>(a, b, r, n); break;
    }
}
extern "C" void fp_host_dot(int field, int t, const u32* a, const u32* b, u32* r, long n) {
    switch (field) {
        case 0: dot_t<Bls12_381_Fr>(t, a, b, r, n); break;
        case 1: dot_t<Bn254_Fr>(t, a, b, r, n); break;
        case 2: dot_t<Jubjub_Fr>(t, a, b, r, n); break;
        case 3: dot_t<Bls12_377_Fr>(t, a, b, r, n); break;
    }
}

extern "C" void fp_host_op(int field, int which, const u32* a, const u32* b, u32* r, unsigned long long alpha, long n) {
    for (long i = 0; i < n; i++) {
        const u32 *ai = a + 8 * i, *bi = b + 8 * i; u32* ri = r + 8 * i;
        switch (field) {
            case 0: op<Bls12_381_Fr>(which, ai, bi, ri, alpha); break;
            case 1: op<Bn254_Fr>(which, ai, bi, ri, alpha); break;
            case 2: op<Jubjub_Fr>(which, ai, bi, ri, alpha); break;
            case 3: op<Bls12_377_Fr>(which, ai, bi, ri, alpha); break;
        }
    }
}
