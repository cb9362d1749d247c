// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).  PARITY UNPINNED at the arkworks boundary.
#include "schemes.hpp"
#include <thread>
#include <stdexcept>
#include <algorithm>

namespace orc {

// ---------------------------------------------------------------------------------- utils.rs
G1 inner_product_g(const std::vector<G1>& a, const std::vector<Fr>& b) {  // utils.rs:16-19
    size_t n = std::min(a.size(), b.size());
    G1 acc = G1::identity();                 // Sum for Projective starts at zero
    for (size_t i = 0; i < n; ++i) acc += a[i].mul(b[i]);
    return acc;
}

G1 inner_product_g_mt(const G1* a, const Fr* b, size_t n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<G1> part(nthreads, G1::identity());
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            size_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
            G1 acc = G1::identity();
            for (size_t i = lo; i < hi; ++i) acc += a[i].mul(b[i]);
            part[t] = acc;
        });
    for (auto& x : th) x.join();
    G1 acc = G1::identity();
    for (auto& p : part) acc += p;
    return acc;
}

Fr inner_product_f(const std::vector<Fr>& a, const std::vector<Fr>& b) {  // utils.rs:16-19
    size_t n = std::min(a.size(), b.size());
    Fr acc = Fr::zero();
    for (size_t i = 0; i < n; ++i) acc += a[i] * b[i];
    return acc;
}

std::vector<Fr> powers_of(const Fr& a, size_t n) {  // utils.rs:44-55
    std::vector<Fr> res;
    Fr cur = Fr::one();
    res.push_back(cur);
    for (size_t i = 1; i < n; ++i) {
        cur = cur * a;
        res.push_back(cur);
    }
    return res;
}

std::vector<Fr> invert_domain_at(const Fr& t, size_t N) {  // utils.rs:57-62
    std::vector<Fr> res;
    for (uint64_t i = 0; i < N; ++i) res.push_back(t - Fr::from_u64(i));
    batch_inversion(res);
    return res;
}

// vec_add_and_distribute utils.rs:31-38 : res_i = a_i + x*b_i
static std::vector<Fr> vadd_dist_f(const std::vector<Fr>& a, const std::vector<Fr>& b, const Fr& x) {
    if (a.size() != b.size()) throw std::runtime_error("vec_add_and_distribute: length mismatch");
    std::vector<Fr> r(a.size());
    for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i] * x;
    return r;
}
static std::vector<G1> vadd_dist_g(const std::vector<G1>& a, const std::vector<G1>& b, const Fr& x) {
    if (a.size() != b.size()) throw std::runtime_error("vec_add_and_distribute: length mismatch");
    std::vector<G1> r(a.size());
    for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i].mul(x);
    return r;
}
template <class T>
static void split(const std::vector<T>& a, std::vector<T>& l, std::vector<T>& r) {  // utils.rs:40-42
    l.assign(a.begin(), a.begin() + a.size() / 2);
    r.assign(a.begin() + a.size() / 2, a.end());
}

// Bucket-method MSM for big expected values (not the reference algorithm).
G1 msm_pippenger(const G1Affine* bases, const Fr* scalars, size_t n, int nthreads) {
    const int c = n < 32 ? 3 : (n < 4096 ? 8 : 12);
    const int W = (254 + c - 1) / c;
    std::vector<U256> ks(n);
    for (size_t i = 0; i < n; ++i) ks[i] = scalars[i].to_canonical();
    std::vector<G1> wsum(W, G1::identity());
    auto work = [&](int w) {
        std::vector<G1> buckets((size_t)1 << c, G1::identity());
        for (size_t i = 0; i < n; ++i) {
            uint32_t d = 0;
            for (int b = 0; b < c; ++b) {
                int bit = w * c + b;
                if (bit < 256) d |= (uint32_t)u256_bit(ks[i], bit) << b;
            }
            if (d && !bases[i].infinity) buckets[d] += G1::from_affine(bases[i]);
        }
        G1 run = G1::identity(), acc = G1::identity();
        for (size_t d = ((size_t)1 << c) - 1; d >= 1; --d) {
            run += buckets[d];
            acc += run;
        }
        wsum[w] = acc;
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            for (int w = t; w < W; w += nthreads) work(w);
        });
    for (auto& x : th) x.join();
    G1 acc = G1::identity();
    for (int w = W - 1; w >= 0; --w) {
        for (int b = 0; b < c; ++b) acc = acc.dbl();
        acc += wsum[w];
    }
    return acc;
}

// ---------------------------------------------------------------------------------- precompute.rs
Precompute::Precompute(size_t size_) : size(size_) {  // precompute.rs:25-34
    domain_size = next_pow2(size);
    group_gen = domain_group_gen(domain_size);
    // compute_vanishing_evaluations precompute.rs:46-58
    vanishing_evaluations.assign(size, Fr::zero());
    vanishing_evaluations_inv.assign(size, Fr::zero());
    Fr n_f = Fr::from_u64(size);
    for (size_t i = 0; i < size; ++i) {
        vanishing_evaluations[i] = n_f * group_gen.pow_u64(i).inverse();
        vanishing_evaluations_inv[i] = vanishing_evaluations[i];
    }
    batch_inversion(vanishing_evaluations_inv);
}

std::vector<Fr> Precompute::compute_barycentric_coefficients(const Fr& point) const {  // precompute.rs:72-90
    std::vector<Fr> res(size, Fr::zero());
    if (point.cmp(Fr::from_u64(size)) < 0) {  // strict (quirk Q3)
        res[to_usize(point)] = Fr::one();
        return res;
    }
    Fr t = (point.pow_u64(size) - Fr::one()) * Fr::from_u64(size).inverse();
    for (size_t i = 0; i < size; ++i) {
        Fr pw = group_gen.pow_u64(i);
        res[i] = (t * pw) * (point - pw).inverse();
    }
    return res;
}

// ---------------------------------------------------------------------------------- lagrange_basis.rs
LagrangeBasis LagrangeBasis::from_vec(const std::vector<Fr>& data, uint64_t domain_n) {
    LagrangeBasis lb;
    lb.evals = data;
    lb.max = data.size();
    lb.domain_size = next_pow2(domain_n ? domain_n : data.size());
    lb.group_gen = domain_group_gen(lb.domain_size);
    return lb;
}
LagrangeBasis LagrangeBasis::new_zero(size_t size) {
    return from_vec(std::vector<Fr>(size, Fr::zero()));
}

Fr LagrangeBasis::evaluate(const Precompute& pc, const Fr& point) const {  // lagrange_basis.rs:63-72
    if (point.cmp(Fr::from_u64(max_index())) <= 0) return evals.at(to_usize(point));
    if (point.cmp(Fr::from_u64(domain_size)) <= 0) return Fr::zero();  // inclusive (quirk Q2)
    return evaluate_outside_domain(pc, point);
}
Fr LagrangeBasis::evaluate_outside_domain(const Precompute& pc, const Fr& point) const {  // :74-83
    return inner_product_f(evals, pc.compute_barycentric_coefficients(point));
}

std::vector<Fr> LagrangeBasis::divide_by_vanishing(const Precompute& pc, size_t index) const {  // :91-119
    std::vector<Fr> q(domain_size, Fr::zero());
    Fr index_f = index_to_point(index);
    Fr eval = index >= max ? Fr::zero() : evals[index];
    Fr index_vanishing = pc.vanishing_evaluations.at(index);
    for (size_t i = 0; i < domain_size; ++i) {
        if (i == index) continue;
        Fr i_f = index_to_point(i);
        Fr i_eval = i >= max ? Fr::zero() : evals[i];
        Fr sub = i_eval - eval;
        q[i] = sub * (i_f - index_f).inverse();
        q[index] += sub * index_vanishing * pc.vanishing_evaluations_inv.at(i) * (index_f - i_f).inverse();
    }
    return q;
}

std::vector<Fr> LagrangeBasis::divide_by_vanishing_outside_domain(const Precompute& pc, const Fr& point) const {  // :121-142
    std::vector<Fr> q(domain_size, Fr::zero());
    Fr eval = evaluate(pc, point);
    std::vector<Fr> inversions(domain_size);
    for (size_t i = 0; i < domain_size; ++i) inversions[i] = index_to_point(i) - point;
    batch_inversion(inversions);
    for (size_t i = 0; i < domain_size; ++i) {
        Fr i_eval = i >= max ? Fr::zero() : evals[i];
        q[i] = (i_eval - eval) * inversions[i];
    }
    return q;
}

Fr to_data_item(const G1& c) {  // lib.rs:56-67
    if (c.is_zero()) return Fr::zero();
    uint8_t b[32];
    g1_serialize_compressed(c, b);
    return Fr::from_le_bytes_mod_order(b, 32);  // includes the flag bits (quirk Q7)
}

// ---------------------------------------------------------------------------------- ipa/mod.rs
G1 ipa_commit(const IpaKey& key, const LagrangeBasis& data) { return inner_product_g(key.g, data.evals); }

IpaProof low_level_ipa(const std::vector<G1>& gens_in, const G1& q_in, const std::vector<Fr>& a, const std::vector<Fr>& b,
                       const G1& commitment, const Fr& input_point, Transcript* prev) {  // ipa/mod.rs:268-319
    Fr eval = inner_product_f(a, b);
    std::vector<G1> gens(gens_in.begin(), gens_in.begin() + a.size());
    std::vector<Fr> data = a, other = b;
    Transcript local("ipa");
    Transcript& tr = prev ? *prev : local;
    tr.append(commitment, "C");
    tr.append(input_point, "input point");
    tr.append(eval, "output point");
    IpaProof pf;
    Fr ra = tr.digest("w", true);
    G1 q = q_in.mul(ra);
    while (data.size() > 1) {
        std::vector<Fr> data_l, data_r, b_l, b_r;
        std::vector<G1> gens_l, gens_r;
        split(data, data_l, data_r);
        split(gens, gens_l, gens_r);
        split(other, b_l, b_r);
        G1 y_l = inner_product_g(gens_r, data_l) + q.mul(inner_product_f(data_l, b_r));
        G1 y_r = inner_product_g(gens_l, data_r) + q.mul(inner_product_f(data_r, b_l));
        pf.l.push_back(y_l);
        pf.r.push_back(y_r);
        tr.append(y_l, "L");
        tr.append(y_r, "R");
        ra = tr.digest("x", true);
        data = vadd_dist_f(data_l, data_r, ra);
        gens = vadd_dist_g(gens_r, gens_l, ra);
        other = vadd_dist_f(b_r, b_l, ra);
    }
    pf.tip = data[0];
    pf.y = eval;
    return pf;
}

bool low_level_verify_ipa(const std::vector<G1>& gens, const G1& q_in, const std::vector<Fr>& b, const G1& commitment,
                          const Fr& input_point, const IpaProof& proof, Transcript* prev) {  // ipa/mod.rs:321-360
    G1 c = commitment;
    Transcript local("ipa");
    Transcript& tr = prev ? *prev : local;
    tr.append(commitment, "C");
    tr.append(input_point, "input point");
    tr.append(proof.y, "output point");
    Fr ra = tr.digest("w", true);
    std::vector<Fr> coeffs{Fr::one()};
    G1 q = q_in.mul(ra);
    c += q.mul(proof.y);
    for (size_t i = 0; i < proof.l.size(); ++i) {
        tr.append(proof.l[i], "L");
        tr.append(proof.r[i], "R");
        ra = tr.digest("x", true);
        c = proof.l[i] + c.mul(ra) + proof.r[i].mul(ra.sqr());
        std::vector<Fr> nc;
        for (auto& x : coeffs) {
            nc.push_back(x * ra);
            nc.push_back(x);
        }
        coeffs.swap(nc);
    }
    G1 combined_point = inner_product_g(gens, coeffs);
    Fr combined_b = inner_product_f(b, coeffs);
    return c == combined_point.mul(proof.tip) + q.mul(proof.tip * combined_b);
}

IpaProof ipa_prove_point(const IpaKey& key, const G1& commitment, const Fr& point, const LagrangeBasis& data, Transcript* t) {
    std::vector<Fr> b = key.precompute.compute_barycentric_coefficients(point);
    return low_level_ipa(key.g, key.q, data.evals, b, commitment, point, t);
}
bool ipa_verify_point(const IpaKey& key, const G1& commitment, const Fr& point, const IpaProof& proof, Transcript* t) {
    return low_level_verify_ipa(key.g, key.q, key.precompute.compute_barycentric_coefficients(point), commitment, point,
                                proof, t);
}

IpaCommitProof ipa_prove_commitment(const IpaKey& key, const G1& commitment, const LagrangeBasis& data_in) {  // :199-235
    size_t max = data_in.max_index();
    std::vector<Fr> data(data_in.evals.begin(), data_in.evals.begin() + max + 1);
    std::vector<G1> gens(key.g.begin(), key.g.begin() + max + 1);
    IpaCommitProof pf;
    Transcript tr("ipa");
    tr.append(commitment, "C");
    Fr ra = tr.digest("x", true);
    while (data.size() > 1) {
        std::vector<Fr> data_l, data_r;
        std::vector<G1> gens_l, gens_r;
        split(data, data_l, data_r);
        split(gens, gens_l, gens_r);
        G1 y_l = inner_product_g(gens_r, data_l);
        G1 y_r = inner_product_g(gens_l, data_r);
        pf.l.push_back(y_l);
        pf.r.push_back(y_r);
        tr.append(y_l, "L");
        tr.append(y_r, "R");
        ra = tr.digest("x", true);
        data = vadd_dist_f(data_l, data_r, ra);
        gens = vadd_dist_g(gens_r, gens_l, ra);
    }
    pf.tip = data[0];
    return pf;
}

bool ipa_verify_commitment_proof(const IpaKey& key, const G1& commitment, const IpaCommitProof& proof) {  // :238-265
    std::vector<G1> gens(key.g.begin(), key.g.begin() + ((size_t)1 << proof.l.size()));
    G1 c = commitment;
    std::vector<Fr> coeffs{Fr::one()};
    Transcript tr("ipa");
    tr.append(commitment, "C");
    Fr ra = tr.digest("x", true);
    for (size_t i = 0; i < proof.l.size(); ++i) {
        tr.append(proof.l[i], "L");
        tr.append(proof.r[i], "R");
        ra = tr.digest("x", true);
        c = proof.l[i] + c.mul(ra) + proof.r[i].mul(ra.sqr());
        std::vector<Fr> nc;
        for (auto& x : coeffs) {
            nc.push_back(x * ra);
            nc.push_back(x);
        }
        coeffs.swap(nc);
    }
    return c == inner_product_g(gens, coeffs).mul(proof.tip);
}

// ---------------------------------------------------------------------------------- kzg/mod.rs
// ipa/ipa_point_generator.rs:97-109
bool eth_hash_to_curve(const std::vector<uint8_t>& domain, const uint8_t* msg, size_t msg_len, G1Affine& out) {
    std::vector<uint8_t> m(domain);
    m.insert(m.end(), msg, msg + msg_len);
    uint8_t h[32];
    sha256(m.data(), m.size(), h);
    return g1_affine_from_random_bytes(h, out);
}

// ipa/ipa_point_generator.rs:51-70
std::vector<G1Affine> ipa_crs_gen(const std::vector<uint8_t>& seed, size_t num, uint64_t* next_index) {
    std::vector<G1Affine> res;
    uint64_t i = 0;
    while (res.size() < num) {
        uint8_t le[8];
        for (int k = 0; k < 8; ++k) le[k] = (uint8_t)(i >> (8 * k));
        G1Affine pt;
        if (eth_hash_to_curve(seed, le, 8, pt)) res.push_back(pt);
        ++i;
    }
    if (next_index) *next_index = i;
    return res;
}

KzgKey kzg_setup(size_t max_items, const Fr& tau) {
    // gen(max_items): [G * tau^i]  (kzg_point_generator.rs:32-43), then domain.ifft (kzg/mod.rs:119-123):
    // out_j = (1/n) sum_i (tau^i G) w^{-ij}  ==  G * L_j(tau), over the padded domain of size n.
    uint64_t n = next_pow2(max_items);
    Fr w = domain_group_gen(n);
    Fr winv = w.inverse();
    Fr ninv = Fr::from_u64(n).inverse();
    std::vector<Fr> taup = powers_of(tau, max_items);  // inputs beyond max_items are zero-padded by ifft
    std::vector<G1> lag(n);
    G1 G = G1::generator();
    for (uint64_t j = 0; j < n; ++j) {
        Fr wj = winv.pow_u64(j);
        Fr acc = Fr::zero(), cur = Fr::one();
        for (size_t i = 0; i < max_items; ++i) {
            acc += taup[i] * cur;
            cur = cur * wj;
        }
        lag[j] = G.mul(acc * ninv);
    }
    return KzgKey(lag, tau);
}

G1 kzg_commit(const KzgKey& key, const LagrangeBasis& data) { return inner_product_g(key.lagrange_commitments, data.evals); }

KzgProof kzg_prove_point(const KzgKey& key, const Fr& point, const LagrangeBasis& data, bool* ok) {  // kzg/mod.rs:136-154
    *ok = true;
    KzgProof pf;
    Fr evaluation = data.evaluate(key.precompute, point);
    std::vector<Fr> q;
    if (point.cmp(Fr::from_u64(key.size)) <= 0) {
        uint64_t idx = to_usize(point);
        if (idx >= key.precompute.size) {  // reference panics (vanishing_at out of bounds) — quirk Q2
            *ok = false;
            pf.proof = G1::identity();
            pf.y = evaluation;
            return pf;
        }
        q = data.divide_by_vanishing(key.precompute, idx);
    } else {
        q = data.divide_by_vanishing_outside_domain(key.precompute, point);
    }
    pf.proof = inner_product_g(key.lagrange_commitments, q);
    pf.y = evaluation;
    return pf;
}

bool kzg_verify_point_with_tau(const KzgKey& key, const G1& commitment, const Fr& point, const KzgProof& proof) {
    // kzg/mod.rs:165-189:  e(pi, [tau]_2 - [p]_2) == e(C - [y]_1, H)  <=>  (tau - p) * pi == C - y*G
    Fr p = point.cmp(Fr::from_u64(key.size)) < 0 ? key.precompute.group_gen.pow_u64(to_usize(point)) : point;
    G1 lhs = proof.proof.mul(key.tau - p);
    G1 rhs = commitment - G1::generator().mul(proof.y);
    return lhs == rhs;
}

// ---------------------------------------------------------------------------------- multiproof.rs
struct MultiproofCore {
    LagrangeBasis g, h;
    G1 d, e;
    Fr t;
    Transcript tr{"multiproof"};
};

template <class CommitFn>
static void multiproof_core(size_t max_size, const Precompute& pc, const std::vector<ProverQuery>& queries, CommitFn commit,
                            MultiproofCore& out) {  // multiproof.rs:99-176
    Transcript& tr = out.tr;
    for (auto& q : queries) {
        tr.append(q.commit, "C");
        tr.append_usize(q.z, "z");
        tr.append(q.y, "y");
    }
    Fr r = tr.digest("r", true);
    std::vector<Fr> r_pows = powers_of(r, queries.size());
    // scaled queries grouped by evaluation point (multiproof.rs:119-126)
    std::map<uint64_t, std::vector<std::vector<Fr>>> by_point;
    for (size_t k = 0; k < queries.size(); ++k) {
        std::vector<Fr> s = queries[k].data->evals;
        for (auto& x : s) x = x * r_pows[k];
        by_point[queries[k].z].push_back(s);
    }
    // g(x)  (multiproof.rs:129-148)
    LagrangeBasis g = LagrangeBasis::new_zero(max_size);
    for (auto& kv : by_point) {
        LagrangeBasis total = LagrangeBasis::new_zero(max_size);
        for (auto& s : kv.second)
            for (size_t i = 0; i < s.size() && i < total.evals.size(); ++i) total.evals[i] += s[i];
        std::vector<Fr> quo = total.divide_by_vanishing(pc, kv.first);
        for (size_t i = 0; i < quo.size() && i < g.evals.size(); ++i) g.evals[i] += quo[i];
    }
    out.d = commit(g);
    tr.append(out.d, "D");
    out.t = tr.digest("t", true);
    std::vector<Fr> inversions = invert_domain_at(out.t, max_size);
    LagrangeBasis h = LagrangeBasis::new_zero(max_size);  // multiproof.rs:161-166
    for (auto& kv : by_point)
        for (auto& s : kv.second)
            for (size_t i = 0; i < s.size() && i < h.evals.size(); ++i) h.evals[i] += s[i] * inversions[kv.first];
    out.e = commit(h);
    tr.append(out.e, "E");
    out.g = g;
    out.h = h;
}

IpaMultiproof ipa_prove_multiproof(const IpaKey& key, const std::vector<ProverQuery>& queries) {
    MultiproofCore c;
    multiproof_core(key.g.size(), key.precompute, queries, [&](const LagrangeBasis& d) { return ipa_commit(key, d); }, c);
    LagrangeBasis hmg = c.h;
    for (size_t i = 0; i < hmg.evals.size(); ++i) hmg.evals[i] = c.h.evals[i] - c.g.evals[i];
    IpaMultiproof mp;
    mp.d = c.d;
    mp.proof = ipa_prove_point(key, c.e - c.d, c.t, hmg, &c.tr);
    return mp;
}

KzgMultiproof kzg_prove_multiproof(const KzgKey& key, const std::vector<ProverQuery>& queries) {
    MultiproofCore c;
    multiproof_core(key.size, key.precompute, queries, [&](const LagrangeBasis& d) { return kzg_commit(key, d); }, c);
    LagrangeBasis hmg = c.h;
    for (size_t i = 0; i < hmg.evals.size(); ++i) hmg.evals[i] = c.h.evals[i] - c.g.evals[i];
    KzgMultiproof mp;
    mp.d = c.d;
    bool ok;
    mp.proof = kzg_prove_point(key, c.t, hmg, &ok);
    return mp;
}

template <class VerifyFn>
static bool verify_multiproof_core(size_t max_size, const std::vector<VerifierQuery>& queries, const G1& d, VerifyFn verify) {
    // multiproof.rs:178-215
    Transcript tr("multiproof");
    for (auto& q : queries) {
        tr.append(q.commit, "C");
        tr.append_usize(q.z, "z");
        tr.append(q.y, "y");
    }
    Fr r = tr.digest("r", true);
    tr.append(d, "D");
    Fr t = tr.digest("t", true);
    Fr r_pow = Fr::one();
    std::vector<Fr> inversions = invert_domain_at(t, max_size);
    // e_coeffs keyed by commitment (HashMap<&Commitment, F>; result is order independent)
    std::vector<std::pair<G1, Fr>> e_coeffs;
    for (auto& q : queries) {
        Fr e_coeff = r_pow * inversions.at(q.z);
        bool found = false;
        for (auto& kv : e_coeffs)
            if (kv.first == q.commit) {
                kv.second += e_coeff;
                found = true;
                break;
            }
        if (!found) e_coeffs.push_back({q.commit, e_coeff});
        r_pow = r_pow * r;  // g2_of_t is computed and never used by the reference (quirk Q4)
    }
    G1 e = G1::identity();
    for (auto& kv : e_coeffs) e += kv.first.mul(kv.second);
    tr.append(e, "E");
    return verify(e - d, t, tr);
}

bool ipa_verify_multiproof(const IpaKey& key, const std::vector<VerifierQuery>& q, const IpaMultiproof& proof) {
    return verify_multiproof_core(key.g.size(), q, proof.d, [&](const G1& c, const Fr& t, Transcript& tr) {
        return ipa_verify_point(key, c, t, proof.proof, &tr);
    });
}
bool kzg_verify_multiproof_with_tau(const KzgKey& key, const std::vector<VerifierQuery>& q, const KzgMultiproof& proof) {
    return verify_multiproof_core(key.size, q, proof.d, [&](const G1& c, const Fr& t, Transcript&) {
        return kzg_verify_point_with_tau(key, c, t, proof.proof);
    });
}

// ---------------------------------------------------------------------------------- verkle-tree node.rs
static TreeNode* new_extension(const uint8_t* key, size_t key_len, const uint8_t value[32]) {
    TreeNode* n = new TreeNode();
    n->internal = false;
    n->stem.assign(key, key + key_len);
    std::array<uint8_t, 32> v;
    memcpy(v.data(), value, 32);
    n->leaves[key[key_len - 1]] = v;
    return n;
}

static void node_insert(TreeNode* self, const uint8_t* key, size_t N, const uint8_t value[32], size_t cur_depth) {  // node.rs:133-197
    if (!self->internal) {
        if (memcmp(self->stem.data(), key, N) != 0) throw std::runtime_error("Traversed to extension node with differing stem");
        std::array<uint8_t, 32> v;
        memcpy(v.data(), value, 32);
        self->leaves[key[N - 1]] = v;
        return;
    }
    uint32_t k = key[cur_depth];
    auto it = self->children.find(k);
    if (it == self->children.end()) {
        self->children[k] = new_extension(key, N, value);
        return;
    }
    TreeNode* child = it->second;
    if (!child->internal) {
        if (memcmp(child->stem.data(), key, N) == 0 || cur_depth == N - 2) {
            node_insert(child, key, N, value, cur_depth + 1);
        } else {
            // next_diff_depth lib.rs:50-59
            size_t d = cur_depth + 1;
            while (d < N) {
                if (child->stem[d] != key[d]) break;
                ++d;
            }
            TreeNode* in = new TreeNode();
            in->internal = true;
            in->children[key[d]] = new_extension(key, N, value);
            in->children[child->stem[d]] = child;
            self->children[k] = in;
        }
    } else {
        node_insert(child, key, N, value, cur_depth + 1);
    }
}

void Tree::insert(const uint8_t* key, const uint8_t value[32]) { node_insert(&root, key, key_len, value, 0); }

static G1 commit_vec(const std::vector<G1>& bases, const std::vector<Fr>& v) { return inner_product_g(bases, v); }

static G1 node_commit(const TreeNode* n, const std::vector<G1>& bases, size_t W) {  // node.rs:212-277
    if (!n->internal) {
        std::vector<Fr> c1(W, Fr::zero()), c2(W, Fr::zero());
        for (auto& kv : n->leaves) {
            size_t index = kv.first;
            Fr low = Fr::from_le_bytes_mod_order(kv.second.data(), 16);
            Fr high = Fr::from_le_bytes_mod_order(kv.second.data() + 16, 16);
            size_t il = (2 * index) % W, ih = (2 * index + 1) % W;
            if (index < W / 2) {
                c1[il] = low;
                c1[ih] = high;
            } else {
                c2[il] = low;
                c2[ih] = high;
            }
        }
        G1 C1 = commit_vec(bases, c1), C2 = commit_vec(bases, c2);
        std::vector<Fr> ext{Fr::one(), Fr::from_le_bytes_mod_order(n->stem.data(), n->stem.size()), to_data_item(C1),
                            to_data_item(C2)};
        return commit_vec(bases, ext);
    }
    std::vector<Fr> vc(256, Fr::zero());
    for (auto& kv : n->children) vc.at(kv.first) = to_data_item(node_commit(kv.second, bases, W));
    return commit_vec(bases, vc);
}

G1 Tree::commitment(const std::vector<G1>& bases, size_t ext_width) const { return node_commit(&root, bases, ext_width); }

}  // namespace orc
