// poseidon_team.cuh -- latency-oriented two-to-one compression for t = 3: THREE WARPS cooperate on 32 hashes.
//
// The top ~13 levels of a Merkle tree have fewer nodes than the GPU has warp schedulers; with one hash per
// thread each such level costs one full single-warp permutation (~0.23 ms for BLS12-381 Fr), all of it on
// one scheduler.  Here warp w of a 96-thread CTA owns state lane w of 32 hashes, so the three S-boxes and
// the three MDS rows of a full round run on three schedulers at once, and in a partial round warp 0 computes
// the S-box and the new lane 0 while warps 1, 2 apply their column updates.  Lanes are exchanged through a
// double-buffered shared-memory array with ONE __syncthreads per round:
//   phase A (before the barrier): add constants, S-box, publish own lane into buffer r&1
//   phase B (after the barrier):  read the lanes needed, apply the linear layer; in sparse partial rounds
//                                 warps 1, 2 publish their updated lane into buffer (r+1)&1 for the next round.
// Buffer (r+1)&1 was last read in phase B of round r-1, which every warp has left before the barrier of round r,
// so the early publish cannot race.  Same schedule, same arithmetic bodies as poseidon.cuh => identical outputs.
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
            if (w != 0) st_elem(team_slot(xb, b, w, lane), s);
        }
        if (w == 0) {
            pos_sbox<F>(s, P.alpha, top_bit, pm);
            st_elem(team_slot(xb, b, 0, lane), s);
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
            u32 v[3][8], d[8];
            fp_copy(v[0], s);
            ld_elem(v[1], team_slot(xb, b, 1, lane));
            ld_elem(v[2], team_slot(xb, b, 2, lane));
            fp_dot<F, 3>(d, v, row, pm);
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
            st_elem(team_slot(xb, b ^ 1, w, lane), s);          // for the next round (see header)
        }
    }
}

}  // namespace cpb
