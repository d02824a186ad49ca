// poseidon_kernels.cuh -- Poseidon kernel templates, the context struct and the launch wrappers.
// Included by cpb_poseidon.cu (C-ABI, dispatch) and by the per-field instantiation units
// poseidon_inst_*.cu, which exist only so that nvcc can compile the (field, width) grid in parallel.
#pragma once
#include <mutex>

#include "common.cuh"
#include "poseidon.cuh"
#include "poseidon_host.hpp"
#include "poseidon_team.cuh"

namespace cpb {

constexpr int kBlock = 128;
// resident CTAs per SM the register allocator must allow: 5 for the narrow states (96 registers), 1 for wide ones
#ifndef CPB_POS_MINB3
#define CPB_POS_MINB3 5
#endif
constexpr int pos_min_blocks(int t) { return t <= 3 ? CPB_POS_MINB3 : t <= 5 ? 3 : 1; }

template <class F, int T>
__global__ void __launch_bounds__(kBlock, pos_min_blocks(T))
k_poseidon_crh(PoseidonDev P, const u32* __restrict__ consts, const u32* __restrict__ in, u32* __restrict__ out,
               long n, long len, long n_out) {
    extern __shared__ __align__(16) u32 cs[];
    __shared__ __align__(8) unsigned long long mbar;
    tma_stage_to_smem(cs, consts, (unsigned)P.n_elems * 32u, &mbar);
    const u32* ct = cs + (int)threadIdx.x * P.zero;   // == cs, but not provably warp-uniform
    u32 pm[8];
    ld_elem(pm, ct + 8 * P.off_mod);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        pos_sponge<F, T>(out + 8 * n_out * i, n_out, in + 8 * len * i, len, P, ct, pm);
    }
}

template <class F, int T>
__global__ void __launch_bounds__(kBlock, pos_min_blocks(T))
k_poseidon_permute(PoseidonDev P, const u32* __restrict__ consts, const u32* __restrict__ in, u32* __restrict__ out,
                   long n) {
    extern __shared__ __align__(16) u32 cs[];
    __shared__ __align__(8) unsigned long long mbar;
    tma_stage_to_smem(cs, consts, (unsigned)P.n_elems * 32u, &mbar);
    const u32* ct = cs + (int)threadIdx.x * P.zero;
    u32 pm[8];
    ld_elem(pm, ct + 8 * P.off_mod);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 s[T][8];
#pragma unroll
        for (int j = 0; j < T; j++) ld_elem(s[j], in + 8 * (T * i + j));
        pos_permute<F, T>(s, P, ct, pm);
#pragma unroll
        for (int j = 0; j < T; j++) st_elem(out + 8 * (T * i + j), s[j]);
    }
}

}  // namespace cpb

struct cpb_poseidon_ctx {
    typedef cpb::u32 u32;
    int field_id = 0, device = 0, sms = 148;
    cpb::host::PoseidonSchedule sched;
    cpb::PoseidonDev dev{};
    u32* d_consts = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    cpb::Scratch s_in, s_out, s_aux;
    std::mutex side_mu;                 // guards creation of side[] only (mu may already be held by a host-pointer call)
    cudaStream_t side[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // subtree streams of the Merkle build
};


namespace cpb {

// cudaFuncAttributeMaxDynamicSharedMemorySize and the occupancy are properties of (kernel, device): the cache key holds the
// current device, so one host thread driving contexts on several GPUs re-configures the kernel on each of them.
template <class K> cpb_status configure_kernel(K kernel, size_t smem, int block, int& occ_out) {
    static thread_local const void* last = nullptr;
    static thread_local int last_occ = 0, last_dev = -1;
    static thread_local size_t last_smem = 0;
    int dev = -1;
    CPB_CUDA(cudaGetDevice(&dev));
    if (last != (const void*)kernel || last_smem != smem || last_dev != dev) {
        if (smem > 48 * 1024) CPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem));
        if (occ < 1) return fail(CPB_CUDA_ERROR, "kernel does not fit on an SM (smem=%zu)", smem);
        last = (const void*)kernel; last_occ = occ; last_smem = smem; last_dev = dev;
    }
    occ_out = last_occ;
    return CPB_OK;
}

template <class K> cpb_status grid_for(K kernel, size_t smem, int sms, long n, int& grid) {
    int last_occ = 0;
    CPB_TRY(configure_kernel(kernel, smem, kBlock, last_occ));
    // One CTA per 128 hashes (capped; the kernels grid-stride).  Short-lived CTAs instead of one persistent
    // wave let the block scheduler interleave kernels from different streams: the latency-bound top
    // levels of one Merkle subtree then overlap with the bulk hashing of the next (merkle build).
    long need = (n + kBlock - 1) / kBlock;
    long cap = (long)sms * last_occ * 64;
    grid = (int)(need < cap ? need : cap);
    if (grid < 1) grid = 1;
    return CPB_OK;
}

template <class F, int T>
cpb_status launch_crh_ft(cpb_poseidon_ctx* c, const u32* in, size_t len, u32* out, size_t n_out, size_t n, cudaStream_t st) {
    size_t smem = (size_t)c->dev.n_elems * 32;
    int grid = 1;
    CPB_TRY(grid_for(k_poseidon_crh<F, T>, smem, c->sms, (long)n, grid));
    k_poseidon_crh<F, T><<<grid, kBlock, smem, st>>>(c->dev, c->d_consts, in, out, (long)n, (long)len, (long)n_out);
    CPB_CUDA(cudaGetLastError());
    return CPB_OK;
}
template <class F, int T>
cpb_status launch_permute_ft(cpb_poseidon_ctx* c, const u32* in, u32* out, size_t n, cudaStream_t st) {
    size_t smem = (size_t)c->dev.n_elems * 32;
    int grid = 1;
    CPB_TRY(grid_for(k_poseidon_permute<F, T>, smem, c->sms, (long)n, grid));
    k_poseidon_permute<F, T><<<grid, kBlock, smem, st>>>(c->dev, c->d_consts, in, out, (long)n);
    CPB_CUDA(cudaGetLastError());
    return CPB_OK;
}


// One Path::verify per thread; both round schedules staged in shared memory.
template <class F, int T>
__global__ void __launch_bounds__(kBlock, pos_min_blocks(T))
k_poseidon_verify_paths(PoseidonDev PL, const u32* __restrict__ consts_l, PoseidonDev PN, const u32* __restrict__ consts_n,
                        const u32* __restrict__ root, const u32* __restrict__ leaves, long leaf_len,
                        const u32* __restrict__ siblings, const u32* __restrict__ paths, int plen,
                        const unsigned long long* __restrict__ indexes, unsigned char* __restrict__ ok, long n) {
    extern __shared__ __align__(16) u32 cs[];
    __shared__ __align__(8) unsigned long long mbar[2];
    u32* csn = cs + 8 * PL.n_elems;
    tma_stage_to_smem(cs, consts_l, (unsigned)PL.n_elems * 32u, &mbar[0]);
    tma_stage_to_smem(csn, consts_n, (unsigned)PN.n_elems * 32u, &mbar[1]);
    const int z = (int)threadIdx.x * PL.zero;
    u32 pm[8];
    ld_elem(pm, cs + z + 8 * PL.off_mod);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        ok[i] = pos_verify_path<F, T>(leaves + 8 * leaf_len * i, leaf_len, siblings + 8 * i, paths + 8 * (long)plen * i, plen, indexes[i],
                                      root, PL, cs + z, PN, csn + z, pm) ? 1 : 0;
}

template <class F, int T>
cpb_status launch_verify_ft(cpb_poseidon_ctx* leaf, cpb_poseidon_ctx* node, const u32* root, const u32* leaves, size_t leaf_len,
                            const u32* siblings, const u32* paths, int plen, const unsigned long long* indexes,
                            unsigned char* ok, size_t n, cudaStream_t st) {
    size_t smem = ((size_t)leaf->dev.n_elems + node->dev.n_elems) * 32;
    if (smem > 200 * 1024) return fail(CPB_UNSUPPORTED, "round schedules (%zu B) exceed shared memory", smem);
    int grid = 1;
    CPB_TRY(grid_for(k_poseidon_verify_paths<F, T>, smem, leaf->sms, (long)n, grid));
    k_poseidon_verify_paths<F, T><<<grid, kBlock, smem, st>>>(leaf->dev, leaf->d_consts, node->dev, node->d_consts, root, leaves,
                                                              (long)leaf_len, siblings, paths, plen, indexes, ok, (long)n);
    CPB_CUDA(cudaGetLastError());
    return CPB_OK;
}

// Three warps per 32 two-to-one hashes (poseidon_team.cuh); t = 3, capacity 1 only.
template <class F>
__global__ void __launch_bounds__(96)
k_poseidon_compress_team(PoseidonDev P, const u32* __restrict__ consts, const u32* __restrict__ pairs, u32* __restrict__ out, long n) {
    extern __shared__ __align__(16) u32 cs[];
    __shared__ __align__(8) unsigned long long mbar;
    tma_stage_to_smem(cs, consts, (unsigned)P.n_elems * 32u, &mbar);
    u32* xb = cs + 8 * P.n_elems;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32* ct = cs + (int)threadIdx.x * P.zero;
    u32 pm[8];
    ld_elem(pm, ct + 8 * P.off_mod);
    int top_bit = 0;
    for (int i = 63; i > 0; i--)
        if ((P.alpha >> i) & 1) { top_bit = i; break; }
    const int total = P.rf + P.rp;
    const long nblk = (n + 31) / 32;
    for (long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long i = blk * 32 + lane;
        const bool active = i < n;
        const long ii = active ? i : n - 1;
        u32 s[8];
        if (w == 0) fp_zero(s);
        else ld_elem(s, pairs + 16 * ii + 8 * (w - 1));
#pragma unroll 1
        for (int r = 0; r < total; r++) {
            team_phase_a<F>(s, w, lane, r, P, ct, pm, xb, top_bit);
            __syncthreads();
            team_phase_b<F>(s, w, lane, r, P, ct, pm, xb);
        }
        if (w == 1 && active) st_elem(out + 8 * i, s);
        __syncthreads();                                   // the exchange buffers are reused by the next 32 hashes
    }
}

template <class F>
cpb_status launch_team_f(cpb_poseidon_ctx* c, const u32* pairs, u32* out, size_t n, cudaStream_t st) {
    size_t smem = (size_t)c->dev.n_elems * 32 + (size_t)kTeamXbWords * 4;
    int occ = 0;
    CPB_TRY(configure_kernel(k_poseidon_compress_team<F>, smem, 96, occ));
    long nblk = ((long)n + 31) / 32;
    k_poseidon_compress_team<F><<<(int)nblk, 96, smem, st>>>(c->dev, c->d_consts, pairs, out, (long)n);
    CPB_CUDA(cudaGetLastError());
    return CPB_OK;
}

// explicit instantiations live in poseidon_inst_<field>.cu
#define CPB_POS_WIDTHS(M, F) M(F, 2) M(F, 3) M(F, 4) M(F, 5) M(F, 6) M(F, 7) M(F, 8) M(F, 9)
#define CPB_POS_INSTANTIATE(F, T)                                                                                        \
    template cpb_status launch_crh_ft<F, T>(cpb_poseidon_ctx*, const u32*, size_t, u32*, size_t, size_t, cudaStream_t);            \
    template cpb_status launch_permute_ft<F, T>(cpb_poseidon_ctx*, const u32*, u32*, size_t, cudaStream_t);             \
    template cpb_status launch_verify_ft<F, T>(cpb_poseidon_ctx*, cpb_poseidon_ctx*, const u32*, const u32*, size_t, const u32*, \
                                               const u32*, int, const unsigned long long*, unsigned char*, size_t, cudaStream_t);
#define CPB_POS_INSTANTIATE_TEAM(F) template cpb_status launch_team_f<F>(cpb_poseidon_ctx*, const u32*, u32*, size_t, cudaStream_t);
#define CPB_POS_EXTERN_TEAM(F) extern template cpb_status launch_team_f<F>(cpb_poseidon_ctx*, const u32*, u32*, size_t, cudaStream_t);
#define CPB_POS_EXTERN(F, T)                                                                                             \
    extern template cpb_status launch_crh_ft<F, T>(cpb_poseidon_ctx*, const u32*, size_t, u32*, size_t, size_t, cudaStream_t);     \
    extern template cpb_status launch_permute_ft<F, T>(cpb_poseidon_ctx*, const u32*, u32*, size_t, cudaStream_t);      \
    extern template cpb_status launch_verify_ft<F, T>(cpb_poseidon_ctx*, cpb_poseidon_ctx*, const u32*, const u32*, size_t, const u32*, \
                                                      const u32*, int, const unsigned long long*, unsigned char*, size_t, cudaStream_t);

}  // namespace cpb
