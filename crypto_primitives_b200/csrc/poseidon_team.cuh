// poseidon_team.cuh -- latency-oriented two-to-one compression for t = 3: THREE WARPS cooperate on 32 hashes.
//
// The top ~13 levels of a Merkle tree have fewer nodes than the GPU has warp schedulers; with one hash per
// thread each such level costs one full single-warp permutation (~0.23 ms for BLS12-381 Fr), all of it on
// one scheduler.  Here warp w of a 96-thread CTA owns state lane w of 32 hashes, so the three S-boxes and
// the three MDS rows of a full round run on three schedulers at once.  In a sparse partial round warp 0 computes
// the S-box while warps 1, 2 -- which have nothing else to do -- form their terms w_j * s_j of the new lane 0
// (a field product is a field product: the sum of separately reduced terms is bit-identical to the lazy dot
// product of poseidon.cuh); after the barrier warp 0 needs one multiplication and two additions instead of a
// three-term dot product, and warps 1, 2 apply their column updates.  Lanes are exchanged through a
// double-buffered shared-memory array with ONE __syncthreads per round:
//   phase A (before the barrier): add constants, S-box (or the dot-product term), publish into buffer r&1
//   phase B (after the barrier):  read what is needed, apply the linear layer.
// Buffer r&1 was last read in phase B of round r-2, which every warp has left before the barrier of round r-1,
// so a publish in phase A of round r cannot race.  Same schedule and field values as poseidon.cuh => identical outputs.
// The phase functions are CPB_HD: tests/host runs them for w = 0, 1, 2 in sequence as the CPU model of the kernel.
#pragma once
#include "poseidon.cuh"

namespace cpb {

// exchange buffer: [2 buffers][3 state lanes][32 hashes] elements of 8 words
CPB_HD u32* team_slot(u32* xb, int buf, int w, int lane) { return xb + (((buf * 3 + w) * 32 + lane) * 8); }
constexpr int kTeamXbWords = 2 * 3 * 32 * 8;

template <class F>
CPB_HD void team_phase_a(u32* s, int w, int lane, int r, const PoseidonDev& P, const u32* cs, const u32* pm, u32* xb, int top_bit) {
    const int half = P.rf / 2;
    const bool full = r < half || r >= half + P.rp;
    const int k = r - half, b = r & 1;
    u32 c[8];
    if (full) {
        const int fr = r < half ? r : r - P.rp;
        ld_elem(c, cs + 8 * (P.off_c + fr * 3 + w));
        fp_add<F>(s, s, c);
        pos_sbox<F>(s, P.alpha, top_bit, pm);
        st_elem(team_slot(xb, b, w, lane), s);
    } else if (!P.sparse) {
        ld_elem(c, cs + 8 * (P.off_arkp + k * 3 + w));
        fp_add<F>(s, s, c);
        if (w == 0) pos_sbox<F>(s, P.alpha, top_bit, pm);
        st_elem(team_slot(xb, b, w, lane), s);
    } else {
        if (k == 0) {
            ld_elem(c, cs + 8 * (P.off_cp0 + w));
            fp_add<F>(s, s, c);
        }
        if (w == 0) {
            pos_sbox<F>(s, P.alpha, top_bit, pm);
            st_elem(team_slot(xb, b, 0, lane), s);
        } else {
            u32 u[8];
            ld_elem(c, cs + 8 * (P.off_sp + k * 5 + w));       // row = [m00, w1, w2 | v1, v2]
            fp_mul<F>(u, s, c, pm);
            st_elem(team_slot(xb, b, w, lane), u);              // the term w_j * s_j, not the lane
        }
    }
}

template <class F>
CPB_HD void team_phase_b(u32* s, int w, int lane, int r, const PoseidonDev& P, const u32* cs, const u32* pm, u32* xb) {
    const int half = P.rf / 2;
    const bool full = r < half || r >= half + P.rp;
    const int k = r - half, b = r & 1;
    if (full || !P.sparse) {
        u32 v[3][8];
#pragma unroll
        for (int j = 0; j < 3; j++) ld_elem(v[j], team_slot(xb, b, j, lane));
        const u32* M = cs + 8 * ((full && r == half - 1) ? P.off_mpre : P.off_m);
        u32 d[8];
        fp_dot<F, 3>(d, v, M + 8 * 3 * w, pm);
        fp_copy(s, d);
    } else {
        const u32* row = cs + 8 * (P.off_sp + k * 5);          // [m00, w1, w2 | v1, v2]
        if (w == 0) {
            u32 u1[8], u2[8], m00[8], d[8];
            ld_elem(m00, row);
            ld_elem(u1, team_slot(xb, b, 1, lane));
            ld_elem(u2, team_slot(xb, b, 2, lane));
            fp_mul<F>(d, s, m00, pm);
            fp_add<F>(d, d, u1);
            fp_add<F>(d, d, u2);
            if (k + 1 < P.rp) {
                u32 c[8];
                ld_elem(c, cs + 8 * (P.off_pc + k + 1));
                fp_add<F>(s, d, c);
            } else {
                fp_copy(s, d);
            }
        } else {
            u32 s0[8], c[8], tmp[8];
            ld_elem(s0, team_slot(xb, b, 0, lane));
            ld_elem(c, row + 8 * (3 + (w - 1)));
            fp_mul<F>(tmp, s0, c, pm);
            fp_add<F>(s, s, tmp);
        }
    }
}

}  // namespace cpb
