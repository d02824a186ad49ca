// cpb200.hpp -- header-only C++ mirror of the reference's scheme interface over the C-ABI (cpb200.h).
//
// The reference is Rust; where no Rust toolchain exists this is the compiled-language host side:
// the same associated-function shape as the traits it mirrors (parameters first, stateless),
//   CRHScheme / TwoToOneCRHScheme   R/crh/mod.rs:18-51
//   CommitmentScheme                R/commitment/mod.rs:15-27
//   merkle_tree::MerkleTree         R/merkle_tree/mod.rs:381-533
// with batch entry points beside the single-shot ones (a GPU is amortised only over a batch).
// Errors: the reference's panics / Err become cpb::Error (std::runtime_error with the status).
// Elements are cpb::Fe = 4 x uint64_t Montgomery limbs (the memory image of ark-ff's Fp256).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "cpb200.h"

namespace cpb {

using Fe = std::array<uint64_t, 4>;
struct Affine { Fe x, y; };

struct Error : std::runtime_error {
    cpb_status status;
    Error(cpb_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(cpb_status s) {
    if (s != CPB_OK) throw Error(s, std::string("cpb status ") + std::to_string((int)s) + ": " + cpb_last_error());
}

namespace poseidon {

// PoseidonConfig<F> (R/sponge/poseidon/mod.rs:26-45) + the device context built from it.
class Config {
public:
    int field_id, full_rounds, partial_rounds, rate, capacity;
    uint64_t alpha;
    std::vector<Fe> ark, mds;   // ark[round * t + i], mds[i * t + j]

    // PoseidonConfig::new (mod.rs:189-217)
    Config(int field, int full_rounds_, int partial_rounds_, uint64_t alpha_, std::vector<Fe> mds_, std::vector<Fe> ark_, int rate_,
           int capacity_, int device = 0)
        : field_id(field), full_rounds(full_rounds_), partial_rounds(partial_rounds_), rate(rate_), capacity(capacity_), alpha(alpha_),
          ark(std::move(ark_)), mds(std::move(mds_)) {
        const size_t t = (size_t)rate + capacity;
        if (ark.size() != (size_t)(full_rounds + partial_rounds) * t || mds.size() != t * t)
            throw Error(CPB_BAD_PARAMS, "PoseidonConfig::new: ark/mds shape");
        cpb_poseidon_ctx* c = nullptr;
        check(cpb_poseidon_ctx_create(field, rate, capacity, full_rounds, partial_rounds, alpha, ark[0].data(), mds[0].data(), device, &c));
        ctx_.reset(c, cpb_poseidon_ctx_destroy);
    }
    // PoseidonDefaultConfigField::get_default_poseidon_parameters (traits.rs:59-103); nullptr where the reference returns None.
    static std::unique_ptr<Config> get_default_poseidon_parameters(int field, int rate, bool optimized_for_weights, int device = 0) {
        uint64_t alpha; int rf, rp, skip;
        if (cpb_poseidon_default_entry(field, rate, optimized_for_weights, &alpha, &rf, &rp, &skip) != CPB_OK) return nullptr;
        uint64_t mod[4];
        check(cpb_field_modulus(field, mod));
        uint64_t bits = 256;
        while (bits && !((mod[(bits - 1) / 64] >> ((bits - 1) % 64)) & 1)) bits--;
        const size_t t = (size_t)rate + 1;
        std::vector<Fe> ark((size_t)(rf + rp) * t), mds(t * t);
        check(cpb_poseidon_find_ark_and_mds(field, bits, rate, rf, rp, skip, ark[0].data(), mds[0].data()));
        return std::unique_ptr<Config>(new Config(field, rf, rp, alpha, std::move(mds), std::move(ark), rate, 1, device));
    }
    cpb_poseidon_ctx* ctx() const { return ctx_.get(); }

private:
    std::shared_ptr<cpb_poseidon_ctx> ctx_;
};

// crh::poseidon::CRH (R/crh/poseidon/mod.rs:15-41)
struct CRH {
    using Parameters = Config;
    static Fe evaluate(const Parameters& p, const std::vector<Fe>& input) {
        Fe out;
        check(cpb_poseidon_crh_batch(p.ctx(), input.empty() ? nullptr : input[0].data(), input.size(), out.data(), 1));
        return out;
    }
    // n inputs of `len` elements each, contiguous
    static std::vector<Fe> evaluate_batch(const Parameters& p, const std::vector<Fe>& inputs, size_t len) {
        const size_t n = len ? inputs.size() / len : 0;
        std::vector<Fe> out(n);
        if (n) check(cpb_poseidon_crh_batch(p.ctx(), inputs[0].data(), len, out[0].data(), n));
        return out;
    }
};

// crh::poseidon::TwoToOneCRH (mod.rs:43-80); evaluate == compress (:58-64)
struct TwoToOneCRH {
    using Parameters = Config;
    static Fe compress(const Parameters& p, const Fe& left, const Fe& right) {
        Fe pair[2] = {left, right}, out;
        check(cpb_poseidon_compress_batch(p.ctx(), pair[0].data(), out.data(), 1));
        return out;
    }
    static Fe evaluate(const Parameters& p, const Fe& left, const Fe& right) { return compress(p, left, right); }
    static std::vector<Fe> compress_batch(const Parameters& p, const std::vector<Fe>& pairs) {
        std::vector<Fe> out(pairs.size() / 2);
        if (!out.empty()) check(cpb_poseidon_compress_batch(p.ctx(), pairs[0].data(), out[0].data(), out.size()));
        return out;
    }
};

}  // namespace poseidon

namespace pedersen {

struct Window { int WINDOW_SIZE, NUM_WINDOWS; };   // R/crh/pedersen/mod.rs:23-26

// pedersen::Parameters (mod.rs:28-31) [+ randomness_generator, R/commitment/pedersen/mod.rs:17-21]
class Parameters {
public:
    Window window;
    Parameters(int curve, Window w, const std::vector<Affine>& generators, const std::vector<Affine>& randomness_generator = {}, int device = 0)
        : window(w) {
        if (generators.size() != (size_t)w.WINDOW_SIZE * w.NUM_WINDOWS) throw Error(CPB_BAD_PARAMS, "Incorrect pp size for window params");
        cpb_pedersen_ctx* c = nullptr;
        check(cpb_pedersen_ctx_create(curve, w.WINDOW_SIZE, w.NUM_WINDOWS, generators[0].x.data(), randomness_generator.size(),
                                      randomness_generator.empty() ? nullptr : randomness_generator[0].x.data(), device, &c));
        ctx_.reset(c, cpb_pedersen_ctx_destroy);
    }
    cpb_pedersen_ctx* ctx() const { return ctx_.get(); }

private:
    std::shared_ptr<cpb_pedersen_ctx> ctx_;
};

struct CRH {   // mod.rs:58-130
    static Affine evaluate(const Parameters& p, const std::vector<uint8_t>& input) {
        Affine out;
        check(cpb_pedersen_crh_batch(p.ctx(), input.data(), input.size(), input.size(), out.x.data(), 1));
        return out;
    }
    static std::vector<Affine> evaluate_batch(const Parameters& p, const uint8_t* inputs, size_t len, size_t n) {
        std::vector<Affine> out(n);
        if (n) check(cpb_pedersen_crh_batch(p.ctx(), inputs, len, len, out[0].x.data(), n));
        return out;
    }
};
struct TwoToOneCRH {   // mod.rs:132-198
    static Affine compress(const Parameters& p, const Affine& l, const Affine& r) {
        Affine kids[2] = {l, r}, out;
        check(cpb_pedersen_two_to_one_batch(p.ctx(), kids[0].x.data(), out.x.data(), 1));
        return out;
    }
};
struct Commitment {   // R/commitment/pedersen/mod.rs:38-106; randomness = 32-byte LE canonical scalar
    static Affine commit(const Parameters& p, const std::vector<uint8_t>& input, const std::array<uint8_t, 32>& randomness) {
        Affine out;
        check(cpb_pedersen_commit_batch(p.ctx(), input.data(), input.size(), input.size(), randomness.data(), out.x.data(), 1));
        return out;
    }
};

}  // namespace pedersen

// MerkleTree<FieldMTConfig> (R/merkle_tree/mod.rs:381-533; Config of R/merkle_tree/tests/mod.rs:198-206)
class PoseidonMerkleTree {
public:
    std::vector<Fe> leaf_nodes, non_leaf_nodes;   // the reference's two arrays (mod.rs:383-395)

    static PoseidonMerkleTree create(const poseidon::Config& leaf, const poseidon::Config& two_to_one, const std::vector<Fe>& leaves, size_t leaf_len) {
        PoseidonMerkleTree t;
        const size_t n = leaf_len ? leaves.size() / leaf_len : 0;
        t.leaf_nodes.resize(n);
        t.non_leaf_nodes.resize(n ? n - 1 : 0);
        check(cpb_merkle_poseidon_build(leaf.ctx(), two_to_one.ctx(), leaves.empty() ? nullptr : leaves[0].data(), leaf_len, n,
                                        n ? t.leaf_nodes[0].data() : nullptr, n > 1 ? t.non_leaf_nodes[0].data() : nullptr));
        return t;
    }
    Fe root() const { return non_leaf_nodes.at(0); }
    size_t height() const {
        size_t h = 1, n = leaf_nodes.size();
        while (n > 1) { n >>= 1; h++; }
        return h;
    }
    // authentication path of leaf `index`, root side first (compute_auth_path, mod.rs:548-573)
    std::vector<Fe> auth_path(size_t index) const {
        std::vector<Fe> path;
        size_t cur = (index + leaf_nodes.size() - 1 - 1) >> 1;   // parent of the leaf's position in the full tree
        while (cur != 0) {
            path.push_back(non_leaf_nodes[(cur & 1) ? cur + 1 : cur - 1]);
            cur = (cur - 1) >> 1;
        }
        return std::vector<Fe>(path.rbegin(), path.rend());
    }
    Fe leaf_sibling_hash(size_t index) const { return leaf_nodes[index ^ 1]; }

    // n x Path::verify (mod.rs:172-212) against `root` in ONE kernel launch: proofs for leaves `indexes` of this tree
    // (generate_proof for each), `leaves` = the claimed leaves, leaf_len elements each.  ok[i] = 1 when path i recomputes the root.
    std::vector<uint8_t> verify_batch(const poseidon::Config& leaf, const poseidon::Config& two_to_one, const Fe& root,
                                      const std::vector<size_t>& indexes, const std::vector<Fe>& leaves, size_t leaf_len) const {
        const size_t n = indexes.size(), plen = height() - 2;
        std::vector<Fe> sib(n), paths(n * plen);
        std::vector<uint64_t> idx(n);
        for (size_t i = 0; i < n; i++) {
            sib[i] = leaf_sibling_hash(indexes[i]);
            std::vector<Fe> ap = auth_path(indexes[i]);
            for (size_t k = 0; k < plen; k++) paths[i * plen + k] = ap[k];
            idx[i] = indexes[i];
        }
        std::vector<uint8_t> ok(n, 0);
        if (n)
            check(cpb_merkle_poseidon_verify_batch(leaf.ctx(), two_to_one.ctx(), root.data(), leaves[0].data(), leaf_len, sib[0].data(),
                                                   plen ? paths[0].data() : nullptr, plen, idx.data(), ok.data(), n));
        return ok;
    }
};

// A group of GPUs driven by this process (include/cpb200.h, "Merkle tree across several GPUs"): MerkleTree::new with the
// leaves sharded over the devices; the result is the reference's two arrays, identical to a one-GPU build.
class GpuGroup {
public:
    explicit GpuGroup(const std::vector<int>& devices) : devices_(devices) {
        cpb_multi* m = nullptr;
        check(cpb_multi_create((int)devices.size(), devices.data(), &m));
        m_.reset(m, cpb_multi_destroy);
    }
    bool uses_nccl() const { return cpb_multi_uses_nccl(m_.get()) != 0; }
    // leaf[d] / two_to_one[d]: the same parameters, created on devices()[d] (Config's `device` argument)
    PoseidonMerkleTree create(const std::vector<const poseidon::Config*>& leaf, const std::vector<const poseidon::Config*>& two_to_one,
                              const std::vector<Fe>& leaves, size_t leaf_len) const {
        if (leaf.size() != devices_.size() || two_to_one.size() != devices_.size()) throw Error(CPB_BAD_PARAMS, "one context per device");
        std::vector<cpb_poseidon_ctx*> lc, nc;
        for (auto* c : leaf) lc.push_back(c->ctx());
        for (auto* c : two_to_one) nc.push_back(c->ctx());
        PoseidonMerkleTree t;
        const size_t n = leaf_len ? leaves.size() / leaf_len : 0;
        t.leaf_nodes.resize(n);
        t.non_leaf_nodes.resize(n ? n - 1 : 0);
        check(cpb_merkle_poseidon_build_multi(m_.get(), lc.data(), nc.data(), leaves.empty() ? nullptr : leaves[0].data(), leaf_len, n,
                                              n ? t.leaf_nodes[0].data() : nullptr, n > 1 ? t.non_leaf_nodes[0].data() : nullptr));
        return t;
    }
    const std::vector<int>& devices() const { return devices_; }

private:
    std::vector<int> devices_;
    std::shared_ptr<cpb_multi> m_;
};

}  // namespace cpb
