// Multi-GPU build through the plain C ABI (include/cpb200.h), as a Rust shim's FFI would call it:
// MerkleTree::new (R/merkle_tree/mod.rs:411-523) of 2^15 two-element leaves over BLS12-381 Fr on ONE GPU
// (cpb_merkle_poseidon_build) and on ALL GPUs (cpb_multi_create + cpb_merkle_poseidon_build_multi, host arrays in the
// reference's layout): the two results must be identical, node for node.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "cpb200.h"

#define CHECK(call) do { cpb_status s__ = (call); if (s__ != CPB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)s__, cpb_last_error()); return 1; } } while (0)

int main() {
    int ndev = cpb_device_count();
    int G = 1;
    while (2 * G <= ndev && 2 * G <= 8) G *= 2;
    if (G < 2) { printf("needs at least 2 GPUs\n"); return 0; }
    const int rate = 2, t = 3;
    uint64_t alpha; int rf, rp, skip;
    CHECK(cpb_poseidon_default_entry(CPB_BLS12_381_FR, rate, 0, &alpha, &rf, &rp, &skip));
    std::vector<uint64_t> ark((size_t)(rf + rp) * t * 4), mds((size_t)t * t * 4);
    CHECK(cpb_poseidon_find_ark_and_mds(CPB_BLS12_381_FR, 255, rate, rf, rp, skip, ark.data(), mds.data()));
    std::vector<cpb_poseidon_ctx*> ctx(G);
    std::vector<int> devs(G);
    for (int d = 0; d < G; d++) {
        devs[d] = d;
        CHECK(cpb_poseidon_ctx_create(CPB_BLS12_381_FR, rate, 1, rf, rp, alpha, ark.data(), mds.data(), d, &ctx[d]));
    }
    const size_t n = 1 << 15, L = 2;
    std::vector<uint64_t> canon(n * L * 4, 0), leaves(n * L * 4);
    for (size_t i = 0; i < n * L; i++) { canon[4 * i] = i * 0x9E3779B97F4A7C15ull + 1; canon[4 * i + 1] = i; }
    CHECK(cpb_field_to_montgomery(CPB_BLS12_381_FR, 0, canon.data(), leaves.data(), n * L));
    std::vector<uint64_t> ln1(n * 4), nn1((n - 1) * 4), lnG(n * 4, 0), nnG((n - 1) * 4, 0);
    CHECK(cpb_merkle_poseidon_build(ctx[0], ctx[0], leaves.data(), L, n, ln1.data(), nn1.data()));
    cpb_multi* m = nullptr;
    CHECK(cpb_multi_create(G, devs.data(), &m));
    for (int rep = 0; rep < 2; rep++)
        CHECK(cpb_merkle_poseidon_build_multi(m, ctx.data(), ctx.data(), leaves.data(), L, n, lnG.data(), nnG.data()));
    if (memcmp(ln1.data(), lnG.data(), ln1.size() * 8) || memcmp(nn1.data(), nnG.data(), nn1.size() * 8)) {
        fprintf(stderr, "multi-GPU arrays differ from the single-GPU build\n");
        return 1;
    }
    // argument checks of the group call
    if (cpb_merkle_poseidon_build_multi(m, ctx.data(), ctx.data(), leaves.data(), L, 24, lnG.data(), nnG.data()) != CPB_NOT_POW2) return 1;
    if (cpb_merkle_poseidon_build_multi(m, ctx.data(), ctx.data(), leaves.data(), L, (size_t)G, lnG.data(), nnG.data()) != CPB_BAD_PARAMS) return 1;
    cpb_multi_destroy(m);
    for (int d = 0; d < G; d++) cpb_poseidon_ctx_destroy(ctx[d]);
    printf("multi-gpu build ok: %d GPUs (%s root exchange), root limb0 = %016llx\n", G, "fused peer-memory or NCCL", (unsigned long long)nnG[0]);
    return 0;
}
