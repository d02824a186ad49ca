// CPU build of tools/fp52.cuh (an experiment, see its header) (fma_rz emulated exactly with 128-bit integers, range assertions on)
// for tests/test_fp52_host.py.  Not part of the product library.
#define CPB_F52_CHECK 1
#include "../../tools/fp52.cuh"
using namespace cpb;

// inputs / outputs: 256-bit little-endian integers as 8 x u32.  Inputs may be lazily reduced (< 4p); outputs canonical.
// which: 0 a*b/2^260   1 a*a/2^260   2 a+b   3 (a0*b0 + a1*b1 + a2*b2)/2^260 + c   (a = 3 elements, b = 3 elements, c = 1)
//        4 chain: x <- x*b, n_chain times (tests that lazily reduced values stay in range)
template <class F> static void op(int which, const u32* a, const u32* b, const u32* c, u32* out, int chain) {
    double pd[5];
    f52::load_modulus<F>(pd);
    u64 r[5];
    if (which == 3) {
        double ad[3][5], bd[15];
        u64 cx[5];
        for (int t = 0; t < 3; t++) {
            u64 x[5], y[5];
            f52::from_words(x, a + 8 * t);
            f52::from_words(y, b + 8 * t);
            f52::to_dbl5(ad[t], x);
            f52::to_dbl5(bd + 5 * t, y);
        }
        f52::from_words(cx, c);
        f52::dot<F, 3>(r, ad, bd, cx, pd);
    } else {
        u64 x[5], y[5];
        double xd[5], yd[5];
        f52::from_words(x, a);
        f52::from_words(y, b);
        f52::to_dbl5(xd, x);
        f52::to_dbl5(yd, y);
        if (which == 0) f52::mul<F>(r, xd, yd, pd);
        if (which == 1) f52::sqr<F>(r, xd, pd);
        if (which == 2) f52::add<F>(r, x, y);
        if (which == 4) {
            for (int i = 0; i < 5; i++) r[i] = x[i];
            for (int k = 0; k < chain; k++) {
                double rd[5];
                f52::to_dbl5(rd, r);
                if (k & 1) f52::sqr<F>(r, rd, pd);
                else f52::mul<F>(r, rd, yd, pd);
                u64 s[5];
                f52::add<F>(s, r, y);            // keep the additive path in the loop too
                for (int i = 0; i < 5; i++) r[i] = s[i];
            }
        }
    }
    f52::canon<F>(r);
    f52::to_words(out, r);
}

extern "C" void fp52_host_op(int field, int which, const u32* a, const u32* b, const u32* c, u32* out, long n, int chain) {
    const int stride = which == 3 ? 24 : 8;
    for (long i = 0; i < n; i++) {
        const u32 *ai = a + stride * i, *bi = b + stride * i, *ci = c + 8 * i;
        u32* oi = out + 8 * i;
        switch (field) {
            case 0: op<Bls12_381_Fr>(which, ai, bi, ci, oi, chain); break;
            case 1: op<Bn254_Fr>(which, ai, bi, ci, oi, chain); break;
            case 2: op<Jubjub_Fr>(which, ai, bi, ci, oi, chain); break;
            case 3: op<Bls12_377_Fr>(which, ai, bi, ci, oi, chain); break;
        }
    }
}
