// C++ host-mirror test (include/cpb200.hpp): reads like the reference's own tests, runs on the GPU.
//   R/sponge/poseidon/mod.rs:381-404   KAT: CRH::evaluate(default rate-2 params, [0,1,2])
//   R/crh/poseidon/mod.rs:58-79        compress(l, r) == evaluate(l, r) == CRH::evaluate([l, r])
//   R/merkle_tree/tests/mod.rs:208-310 tree build / root / path shape; mod.rs:430-433 power-of-two assert
#include <cstdio>
#include <cstdlib>
#include "cpb200.hpp"
using namespace cpb;

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static std::vector<Fe> to_mont(int field, const std::vector<Fe>& canonical) {
    std::vector<Fe> out(canonical.size());
    check(cpb_field_to_montgomery(field, 0, canonical[0].data(), out[0].data(), canonical.size()));
    return out;
}
static Fe from_mont(int field, const Fe& m) {
    Fe out;
    check(cpb_field_from_montgomery(field, 0, m.data(), out.data(), 1));
    return out;
}

int main() {
    auto params = poseidon::Config::get_default_poseidon_parameters(CPB_BLS12_381_FR, 2, false);
    REQUIRE(params && params->full_rounds == 8 && params->partial_rounds == 31 && params->alpha == 17);
    REQUIRE(!poseidon::Config::get_default_poseidon_parameters(CPB_BLS12_381_FR, 9, false));     // None

    std::vector<Fe> in = to_mont(CPB_BLS12_381_FR, {Fe{0, 0, 0, 0}, Fe{1, 0, 0, 0}, Fe{2, 0, 0, 0}});
    Fe h = from_mont(CPB_BLS12_381_FR, poseidon::CRH::evaluate(*params, in));
    // printed for tests/test_gpu_cpp.py, which compares it with tests/golden/reference_kats.json
    printf("CRH([0,1,2]) = %016llx %016llx %016llx %016llx\n", (unsigned long long)h[3], (unsigned long long)h[2], (unsigned long long)h[1],
           (unsigned long long)h[0]);

    Fe l = in[1], r = in[2];
    Fe c = poseidon::TwoToOneCRH::compress(*params, l, r);
    REQUIRE(c == poseidon::TwoToOneCRH::evaluate(*params, l, r));
    REQUIRE(c == poseidon::CRH::evaluate(*params, {l, r}));

    // 8 leaves x 2 elements
    std::vector<Fe> canon;
    for (uint64_t i = 0; i < 16; i++) canon.push_back(Fe{i * 7 + 3, i, 0, 0});
    std::vector<Fe> leaves = to_mont(CPB_BLS12_381_FR, canon);
    auto tree = PoseidonMerkleTree::create(*params, *params, leaves, 2);
    REQUIRE(tree.height() == 4 && tree.leaf_nodes.size() == 8 && tree.non_leaf_nodes.size() == 7);
    std::vector<Fe> digests = poseidon::CRH::evaluate_batch(*params, leaves, 2);
    REQUIRE(digests == tree.leaf_nodes);
    std::vector<Fe> lvl = poseidon::TwoToOneCRH::compress_batch(*params, digests);          // 4 nodes
    std::vector<Fe> lvl2 = poseidon::TwoToOneCRH::compress_batch(*params, lvl);             // 2 nodes
    REQUIRE(poseidon::TwoToOneCRH::compress(*params, lvl2[0], lvl2[1]) == tree.root());
    REQUIRE(tree.non_leaf_nodes[1] == lvl2[0] && tree.non_leaf_nodes[3] == lvl[0]);          // heap order
    REQUIRE(tree.auth_path(5).size() == tree.height() - 2);
    REQUIRE(tree.auth_path(5)[0] == tree.non_leaf_nodes[1]);                                // sibling of the subtree holding leaf 5

    // all 8 paths verified in one launch; a tampered leaf is the only one rejected (Path::verify, mod.rs:172-212)
    std::vector<size_t> all = {0, 1, 2, 3, 4, 5, 6, 7};
    std::vector<uint8_t> ok = tree.verify_batch(*params, *params, tree.root(), all, leaves, 2);
    for (uint8_t v : ok) REQUIRE(v == 1);
    std::vector<Fe> tampered = leaves;
    tampered[2 * 3][0] ^= 1;
    ok = tree.verify_batch(*params, *params, tree.root(), all, tampered, 2);
    for (size_t i = 0; i < 8; i++) REQUIRE(ok[i] == (i == 3 ? 0 : 1));

    bool threw = false;
    try { PoseidonMerkleTree::create(*params, *params, std::vector<Fe>(leaves.begin(), leaves.begin() + 6), 2); }   // 3 leaves
    catch (const Error& e) { threw = e.status == CPB_NOT_POW2; }
    REQUIRE(threw);
    printf("cpp mirror ok\n");
    return 0;
}
