// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).  PARITY UNPINNED at the arkworks boundary.
//
// Line-by-line CPU restatement of the reference's vector-commitment hot path.  Every function
// cites the /root/reference file:line it follows.
#pragma once
#include "bn254.hpp"
#include <map>

namespace orc {

// ---- utils.rs -------------------------------------------------------------------------------
// utils.rs:16-19  inner_product (G x F): one full double-and-add per term, then sum; zip truncates.
G1 inner_product_g(const std::vector<G1>& a, const std::vector<Fr>& b);
// utils.rs:16-19  inner_product (F x F)
Fr inner_product_f(const std::vector<Fr>& a, const std::vector<Fr>& b);
// utils.rs:44-55
std::vector<Fr> powers_of(const Fr& a, size_t n);
// utils.rs:57-62   1/(t - i) for INTEGER i (quirk Q4), batch inverted
std::vector<Fr> invert_domain_at(const Fr& t, size_t N);
// utils.rs:72-74
static inline uint64_t to_usize(const Fr& x) { return x.to_canonical().l[0]; }

// Fast MSM (bucket method) used only to produce large expected values quickly; validated against
// inner_product_g in tests.  NOT the reference algorithm.
G1 msm_pippenger(const G1Affine* bases, const Fr* scalars, size_t n, int nthreads);
// multi-threaded wrapper of the naive algorithm (partial sums per thread, added in order)
G1 inner_product_g_mt(const G1* a, const Fr* b, size_t n, int nthreads);

// ---- precompute.rs --------------------------------------------------------------------------
struct Precompute {  // precompute.rs:11-22
    size_t size;
    uint64_t domain_size;
    Fr group_gen;
    std::vector<Fr> vanishing_evaluations;      // A'(w^i) = n * w^-i      precompute.rs:46-58
    std::vector<Fr> vanishing_evaluations_inv;
    explicit Precompute(size_t size);           // precompute.rs:25-34
    // precompute.rs:72-90
    std::vector<Fr> compute_barycentric_coefficients(const Fr& point) const;
};

// ---- lagrange_basis.rs ----------------------------------------------------------------------
struct LagrangeBasis {  // lagrange_basis.rs:14-21
    std::vector<Fr> evals;
    size_t max;            // == evals.len() at construction (lagrange_basis.rs:25)
    uint64_t domain_size;  // Evaluations' domain size (next pow2 of the domain requested)
    Fr group_gen;
    // lagrange_basis.rs:24-31 from_vec_and_domain / :152-155 from_vec (domain = D::new(len))
    static LagrangeBasis from_vec(const std::vector<Fr>& data, uint64_t domain_n = 0);
    static LagrangeBasis new_zero(size_t size);  // :33-40
    size_t max_index() const { return max - 1; }  // :43-45  max()
    Fr index_to_point(size_t i) const { return group_gen.pow_u64(i); }  // :86-88
    Fr at(size_t i) const { return evals.at(i); }
    Fr evaluate(const Precompute& pc, const Fr& point) const;                  // :63-72
    Fr evaluate_outside_domain(const Precompute& pc, const Fr& point) const;   // :74-83
    std::vector<Fr> divide_by_vanishing(const Precompute& pc, size_t index) const;              // :91-119
    std::vector<Fr> divide_by_vanishing_outside_domain(const Precompute& pc, const Fr& point) const;  // :121-142
};

// lib.rs:56-67 VCCommitment::to_data_item blanket impl
Fr to_data_item(const G1& c);

// ---- ipa/mod.rs -----------------------------------------------------------------------------
struct IpaKey {  // ipa/mod.rs:22-52
    std::vector<G1> g;
    G1 q;
    Precompute precompute;
    IpaKey(const std::vector<G1>& all, size_t N) : g(all.begin(), all.begin() + N), q(all.at(N)), precompute(N) {}
};
struct IpaProof {  // ipa/mod.rs:79-84
    std::vector<G1> l, r;
    Fr tip, y;
};
struct IpaCommitProof {  // ipa/mod.rs:73-77
    std::vector<G1> l, r;
    Fr tip;
};
G1 ipa_commit(const IpaKey& key, const LagrangeBasis& data);  // ipa/mod.rs:130-135
// ipa/mod.rs:268-319
IpaProof low_level_ipa(const std::vector<G1>& gens, const G1& q, const std::vector<Fr>& a, const std::vector<Fr>& b,
                       const G1& commitment, const Fr& input_point, Transcript* prev);
// ipa/mod.rs:321-360
bool low_level_verify_ipa(const std::vector<G1>& gens, const G1& q, const std::vector<Fr>& b, const G1& commitment,
                          const Fr& input_point, const IpaProof& proof, Transcript* prev);
IpaProof ipa_prove_point(const IpaKey& key, const G1& commitment, const Fr& point, const LagrangeBasis& data,
                         Transcript* t);  // ipa/mod.rs:137-154
bool ipa_verify_point(const IpaKey& key, const G1& commitment, const Fr& point, const IpaProof& proof,
                      Transcript* t);  // ipa/mod.rs:165-181
IpaCommitProof ipa_prove_commitment(const IpaKey& key, const G1& commitment, const LagrangeBasis& data);  // :199-235
bool ipa_verify_commitment_proof(const IpaKey& key, const G1& commitment, const IpaCommitProof& proof);   // :238-265

// ---- kzg/mod.rs -----------------------------------------------------------------------------
struct KzgKey {  // kzg/mod.rs:27-57 (g2 replaced by the known secret tau: verification is done on the
                 // scalar side because tau is known in tests; pairings are out of scope)
    size_t size;
    std::vector<G1> lagrange_commitments;
    Fr tau;
    Precompute precompute;
    KzgKey(const std::vector<G1>& lag, const Fr& tau_) : size(lag.size()), lagrange_commitments(lag), tau(tau_), precompute(lag.size()) {}
};
struct KzgProof {  // kzg/mod.rs:81-84
    G1 proof;
    Fr y;
};
// kzg_point_generator.rs:32-43 + kzg/mod.rs:115-124 (powers of tau, then group IFFT == Lagrange
// basis polynomials evaluated at tau; computed on the scalar side, exact).
KzgKey kzg_setup(size_t max_items, const Fr& tau);
G1 kzg_commit(const KzgKey& key, const LagrangeBasis& data);  // kzg/mod.rs:126-134
// kzg/mod.rs:136-154.  ok=false when the reference would index out of bounds (quirk Q2: point == size).
KzgProof kzg_prove_point(const KzgKey& key, const Fr& point, const LagrangeBasis& data, bool* ok);
// kzg/mod.rs:165-189 restated WITHOUT pairings using the known tau:  [tau - p] * proof == C - [y]G
bool kzg_verify_point_with_tau(const KzgKey& key, const G1& commitment, const Fr& point, const KzgProof& proof);

// ---- multiproof.rs --------------------------------------------------------------------------
struct ProverQuery {  // multiproof.rs:25-41
    const LagrangeBasis* data;
    G1 commit;
    uint64_t z;
    Fr y;
};
struct VerifierQuery {  // multiproof.rs:43-53
    G1 commit;
    uint64_t z;
    Fr y;
};
struct IpaMultiproof {  // multiproof.rs:55-58
    IpaProof proof;
    G1 d;
};
struct KzgMultiproof {
    KzgProof proof;
    G1 d;
};
// multiproof.rs:99-176
IpaMultiproof ipa_prove_multiproof(const IpaKey& key, const std::vector<ProverQuery>& q);
KzgMultiproof kzg_prove_multiproof(const KzgKey& key, const std::vector<ProverQuery>& q);
// multiproof.rs:178-215
bool ipa_verify_multiproof(const IpaKey& key, const std::vector<VerifierQuery>& q, const IpaMultiproof& proof);
bool kzg_verify_multiproof_with_tau(const KzgKey& key, const std::vector<VerifierQuery>& q, const KzgMultiproof& proof);

// ---- ipa/ipa_point_generator.rs ---------------------------------------------------------------
// EthereumHashToCurve::hash (:97-109): SHA-256(domain || message) -> Affine::from_random_bytes.
bool eth_hash_to_curve(const std::vector<uint8_t>& domain, const uint8_t* msg, size_t msg_len, G1Affine& out);
// IPAPointGenerator::gen (:51-70): the first `num` indices i = 0, 1, 2 ... (as usize -> 8 LE bytes) that hash to a point;
// *next_index (optional) = the first index not consumed.  The caller checks num <= max (OutOfBounds).
std::vector<G1Affine> ipa_crs_gen(const std::vector<uint8_t>& seed, size_t num, uint64_t* next_index);

// ---- verkle-tree/src/node.rs ----------------------------------------------------------------
// Node::gen_commitment (node.rs:212-277) on a tree built from (key, value) pairs with
// Node::insert (node.rs:133-197).  `ext_width` is the const generic N used for the extension
// layout `(2*idx) % N` (quirk Q6); keys are byte strings of length key_len, the stem is the
// WHOLE key (verkle-tree/src/lib.rs:62-68).  Values are 32-byte, split 16/16 LE
// (verkle-tree/src/lib.rs:194-202 test impl of SplittableValue).
struct TreeNode {
    bool internal;
    std::map<uint32_t, TreeNode*> children;                 // Internal
    std::vector<uint8_t> stem;                              // Extension
    std::map<uint32_t, std::array<uint8_t, 32>> leaves;     // Extension
    ~TreeNode() {
        for (auto& c : children) delete c.second;
    }
};
struct Tree {
    size_t key_len;
    TreeNode root;
    explicit Tree(size_t kl) : key_len(kl) { root.internal = true; }
    void insert(const uint8_t* key, const uint8_t value[32]);
    // commit callback = the scheme's commit over the key's bases (KZG or IPA are the same MSM)
    G1 commitment(const std::vector<G1>& bases, size_t ext_width) const;
};

}  // namespace orc
