// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).  PARITY UNPINNED at the arkworks boundary.
//
// Flat C entry points for ctypes.  Layouts (identical to include/vkzg.h so the same numpy buffers
// feed both sides):  Fr/Fq = 32 bytes, little-endian limbs, Montgomery form (R = 2^256);
// G1 affine = x || y (64 bytes), (0,0) encodes the point at infinity.
#include "schemes.hpp"
#include <thread>
#include <chrono>

using namespace orc;

namespace {

G1Affine load_aff(const uint8_t* p) {
    G1Affine a;
    memcpy(a.x.v.l, p, 32);
    memcpy(a.y.v.l, p + 32, 32);
    a.infinity = a.x.is_zero() && a.y.is_zero();
    return a;
}
G1 load_g1(const uint8_t* p) { return G1::from_affine(load_aff(p)); }
void store_aff(uint8_t* p, const G1Affine& a) {
    if (a.infinity) {
        memset(p, 0, 64);
        return;
    }
    memcpy(p, a.x.v.l, 32);
    memcpy(p + 32, a.y.v.l, 32);
}
void store_g1(uint8_t* p, const G1& g) { store_aff(p, g.to_affine()); }
Fr load_fr(const uint8_t* p) {
    Fr f;
    memcpy(f.v.l, p, 32);
    return f;
}
void store_fr(uint8_t* p, const Fr& f) { memcpy(p, f.v.l, 32); }
std::vector<Fr> load_frs(const uint8_t* p, size_t n) {
    std::vector<Fr> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = load_fr(p + 32 * i);
    return v;
}
std::vector<G1> load_g1s(const uint8_t* p, size_t n) {
    std::vector<G1> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = load_g1(p + 64 * i);
    return v;
}
void store_frs(uint8_t* p, const std::vector<Fr>& v) {
    for (size_t i = 0; i < v.size(); ++i) store_fr(p + 32 * i, v[i]);
}
// normalise many Jacobian points with one inversion
void store_g1s(uint8_t* p, const std::vector<G1>& v) {
    std::vector<Fq> z(v.size());
    for (size_t i = 0; i < v.size(); ++i) z[i] = v[i].Z;
    batch_inversion(z);
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].is_zero()) {
            memset(p + 64 * i, 0, 64);
            continue;
        }
        Fq zi2 = z[i].sqr();
        Fq x = v[i].X * zi2, y = v[i].Y * zi2 * z[i];
        memcpy(p + 64 * i, x.v.l, 32);
        memcpy(p + 64 * i + 32, y.v.l, 32);
    }
}
template <class F>
void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads == 1 || n < 2) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            for (size_t i = t; i < n; i += nthreads) f(i);
        });
    for (auto& x : th) x.join();
}
}  // namespace

#define ORC_TRY try {
#define ORC_CATCH                  \
    }                              \
    catch (const std::exception&) { \
        return -1;                 \
    }                              \
    return 0;

extern "C" {

// ------------------------------------------------------------------ field ops (tag: 0 = Fr, 1 = Fq); op: 0 add 1 sub 2 mul 3 inv(a)
int orc_field_op(int tag, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t n) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) {
        if (tag == 0) {
            Fr x, y, r;
            memcpy(x.v.l, a + 32 * i, 32);
            if (b) memcpy(y.v.l, b + 32 * i, 32);
            r = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : x.inverse();
            memcpy(out + 32 * i, r.v.l, 32);
        } else {
            Fq x, y, r;
            memcpy(x.v.l, a + 32 * i, 32);
            if (b) memcpy(y.v.l, b + 32 * i, 32);
            r = op == 0 ? x + y : op == 1 ? x - y : op == 2 ? x * y : x.inverse();
            memcpy(out + 32 * i, r.v.l, 32);
        }
    }
    ORC_CATCH
}

// canonical LE integer (32 bytes, < modulus) <-> Montgomery
int orc_to_mont(int tag, const uint8_t* in, uint8_t* out, uint64_t n) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) {
        U256 x;
        memcpy(x.l, in + 32 * i, 32);
        if (tag == 0) {
            Fr f = Fr::from_canonical(x);
            memcpy(out + 32 * i, f.v.l, 32);
        } else {
            Fq f = Fq::from_canonical(x);
            memcpy(out + 32 * i, f.v.l, 32);
        }
    }
    ORC_CATCH
}
int orc_from_mont(int tag, const uint8_t* in, uint8_t* out, uint64_t n) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) {
        U256 c;
        if (tag == 0) {
            Fr f;
            memcpy(f.v.l, in + 32 * i, 32);
            c = f.to_canonical();
        } else {
            Fq f;
            memcpy(f.v.l, in + 32 * i, 32);
            c = f.to_canonical();
        }
        memcpy(out + 32 * i, c.l, 32);
    }
    ORC_CATCH
}
int orc_fr_from_le_bytes_mod_order(const uint8_t* bytes, uint64_t len, uint8_t* out) {
    ORC_TRY
    store_fr(out, Fr::from_le_bytes_mod_order(bytes, len));
    ORC_CATCH
}

// ------------------------------------------------------------------ group ops
int orc_g1_generator(uint8_t* out) {
    ORC_TRY
    store_g1(out, G1::generator());
    ORC_CATCH
}
int orc_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    ORC_TRY
    store_g1(out, load_g1(a) + load_g1(b));
    ORC_CATCH
}
int orc_g1_neg(const uint8_t* a, uint8_t* out) {
    ORC_TRY
    store_g1(out, load_g1(a).neg());
    ORC_CATCH
}
int orc_g1_mul(const uint8_t* p, const uint8_t* k, uint8_t* out) {
    ORC_TRY
    store_g1(out, load_g1(p).mul(load_fr(k)));
    ORC_CATCH
}
int orc_g1_on_curve(const uint8_t* p) { return g1_on_curve(load_aff(p)) ? 1 : 0; }
int orc_g1_compress(const uint8_t* p, uint8_t* out32, uint64_t n) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) g1_affine_serialize_compressed(load_aff(p + 64 * i), out32 + 32 * i);
    ORC_CATCH
}
// P_i = (k0 + i*k1) * G  for i < n  (cheap seeded test points: one add per point + one batched inversion)
int orc_points_walk(const uint8_t* k0, const uint8_t* k1, uint64_t n, uint8_t* out) {
    ORC_TRY
    G1 cur = G1::generator().mul(load_fr(k0));
    G1 step = G1::generator().mul(load_fr(k1));
    std::vector<G1> v(n);
    for (uint64_t i = 0; i < n; ++i) {
        v[i] = cur;
        cur += step;
    }
    store_g1s(out, v);
    ORC_CATCH
}
// out_i = k_i * G (multi-threaded)
int orc_g1_mul_gen_batch(const uint8_t* k, uint64_t n, int nthreads, uint8_t* out) {
    ORC_TRY
    std::vector<G1> v(n);
    G1 g = G1::generator();
    parallel_for(n, nthreads, [&](size_t i) { v[i] = g.mul(load_fr(k + 32 * i)); });
    store_g1s(out, v);
    ORC_CATCH
}

// ------------------------------------------------------------------ hashing / transcript
int orc_sha256(const uint8_t* msg, uint64_t len, uint8_t* out32) {
    sha256(msg, len, out32);
    return 0;
}
int orc_expand_message_xmd(const uint8_t* msg, uint64_t len, const uint8_t* dst, uint64_t dst_len, uint64_t n,
                           uint64_t z_pad_len, uint8_t* out) {
    ORC_TRY
    auto v = expand_message_xmd(msg, len, dst, dst_len, n, z_pad_len);
    memcpy(out, v.data(), n);
    ORC_CATCH
}
int orc_hash_to_fr(const uint8_t* msg, uint64_t len, const char* dst, uint8_t* out) {
    ORC_TRY
    store_fr(out, hash_to_fr(msg, len, dst));
    ORC_CATCH
}
int orc_domain_gen(uint64_t n, uint8_t* out) {
    ORC_TRY
    store_fr(out, domain_group_gen(next_pow2(n)));
    ORC_CATCH
}
int orc_to_data_item(const uint8_t* pts, uint64_t n, uint8_t* out) {
    ORC_TRY
    for (uint64_t i = 0; i < n; ++i) store_fr(out + 32 * i, to_data_item(load_g1(pts + 64 * i)));
    ORC_CATCH
}

// ------------------------------------------------------------------ MSM / commit   (M1)
// mode 0 = reference-naive (utils.rs:16-19), 1 = bucket method (expected values only)
int orc_msm(const uint8_t* bases, const uint8_t* scalars, uint64_t n, int mode, int nthreads, uint8_t* out) {
    ORC_TRY
    if (mode == 0) {
        std::vector<G1> b = load_g1s(bases, n);
        std::vector<Fr> s = load_frs(scalars, n);
        store_g1(out, inner_product_g_mt(b.data(), s.data(), n, nthreads));
    } else {
        std::vector<G1Affine> b(n);
        for (uint64_t i = 0; i < n; ++i) b[i] = load_aff(bases + 64 * i);
        std::vector<Fr> s = load_frs(scalars, n);
        store_g1(out, msm_pippenger(b.data(), s.data(), n, nthreads));
    }
    ORC_CATCH
}
// B independent commits of width w against nb bases (zip truncation, quirk Q1)
int orc_commit_batch(const uint8_t* bases, uint64_t nb, const uint8_t* scalars, uint64_t w, uint64_t B, int nthreads,
                     uint8_t* out) {
    ORC_TRY
    std::vector<G1> b = load_g1s(bases, nb);
    std::vector<G1> res(B);
    parallel_for(B, nthreads, [&](size_t k) { res[k] = inner_product_g(b, load_frs(scalars + 32 * w * k, w)); });
    store_g1s(out, res);
    ORC_CATCH
}

// ------------------------------------------------------------------ Lagrange-basis field routines (B1, E1, K1, K2)
int orc_barycentric(uint64_t N, const uint8_t* point, uint8_t* out) {
    ORC_TRY
    Precompute pc(N);
    store_frs(out, pc.compute_barycentric_coefficients(load_fr(point)));
    ORC_CATCH
}
int orc_vanishing(uint64_t N, uint8_t* evals, uint8_t* inv) {
    ORC_TRY
    Precompute pc(N);
    store_frs(evals, pc.vanishing_evaluations);
    store_frs(inv, pc.vanishing_evaluations_inv);
    ORC_CATCH
}
// data: len evaluations over a domain of domain_n; key precompute size N
int orc_evaluate(uint64_t N, const uint8_t* data, uint64_t len, uint64_t domain_n, const uint8_t* point, uint8_t* out) {
    ORC_TRY
    Precompute pc(N);
    LagrangeBasis lb = LagrangeBasis::from_vec(load_frs(data, len), domain_n);
    store_fr(out, lb.evaluate(pc, load_fr(point)));
    ORC_CATCH
}
int orc_divide_by_vanishing(uint64_t N, const uint8_t* data, uint64_t len, uint64_t domain_n, uint64_t index, uint8_t* out) {
    ORC_TRY
    Precompute pc(N);
    LagrangeBasis lb = LagrangeBasis::from_vec(load_frs(data, len), domain_n);
    store_frs(out, lb.divide_by_vanishing(pc, index));
    ORC_CATCH
}
int orc_divide_by_vanishing_outside(uint64_t N, const uint8_t* data, uint64_t len, uint64_t domain_n, const uint8_t* point,
                                    uint8_t* out) {
    ORC_TRY
    Precompute pc(N);
    LagrangeBasis lb = LagrangeBasis::from_vec(load_frs(data, len), domain_n);
    store_frs(out, lb.divide_by_vanishing_outside_domain(pc, load_fr(point)));
    ORC_CATCH
}

// ------------------------------------------------------------------ IPA (I1, I3, I4)
// bases: N+1 affine points (g[0..N], q).  Optional in-flight transcript: (prefix bytes, dst label).
static Transcript make_tr(const uint8_t* prefix, uint64_t prefix_len, const char* dst) {
    Transcript tr(dst ? dst : "ipa");
    if (prefix && prefix_len) tr.state.assign(prefix, prefix + prefix_len);
    return tr;
}
int orc_ipa_prove(const uint8_t* bases, uint64_t N, const uint8_t* a, const uint8_t* commitment, const uint8_t* point,
                  const uint8_t* prefix, uint64_t prefix_len, const char* dst, uint8_t* L, uint8_t* R, uint8_t* tip,
                  uint8_t* y) {
    ORC_TRY
    IpaKey key(load_g1s(bases, N + 1), N);
    LagrangeBasis data = LagrangeBasis::from_vec(load_frs(a, N));
    Transcript tr = make_tr(prefix, prefix_len, dst);
    IpaProof pf = ipa_prove_point(key, load_g1(commitment), load_fr(point), data, &tr);
    store_g1s(L, pf.l);
    store_g1s(R, pf.r);
    store_fr(tip, pf.tip);
    store_fr(y, pf.y);
    ORC_CATCH
}
// B independent proofs (CPU baseline for cfg2): a[B][N], commitments[B], points[B]
int orc_ipa_prove_batch(const uint8_t* bases, uint64_t N, const uint8_t* a, const uint8_t* commitments, const uint8_t* points,
                        uint64_t B, int nthreads, uint8_t* L, uint8_t* R, uint8_t* tip, uint8_t* y) {
    ORC_TRY
    IpaKey key(load_g1s(bases, N + 1), N);
    int lg = 0;
    while ((1ULL << lg) < N) ++lg;
    parallel_for(B, nthreads, [&](size_t k) {
        LagrangeBasis data = LagrangeBasis::from_vec(load_frs(a + 32 * N * k, N));
        IpaProof pf = ipa_prove_point(key, load_g1(commitments + 64 * k), load_fr(points + 32 * k), data, nullptr);
        store_g1s(L + 64 * lg * k, pf.l);
        store_g1s(R + 64 * lg * k, pf.r);
        store_fr(tip + 32 * k, pf.tip);
        store_fr(y + 32 * k, pf.y);
    });
    ORC_CATCH
}
// returns 1 valid, 0 invalid, -1 error
int orc_ipa_verify(const uint8_t* bases, uint64_t N, const uint8_t* commitment, const uint8_t* point, const uint8_t* prefix,
                   uint64_t prefix_len, const char* dst, const uint8_t* L, const uint8_t* R, uint64_t rounds,
                   const uint8_t* tip, const uint8_t* y) {
    try {
        IpaKey key(load_g1s(bases, N + 1), N);
        IpaProof pf;
        pf.l = load_g1s(L, rounds);
        pf.r = load_g1s(R, rounds);
        pf.tip = load_fr(tip);
        pf.y = load_fr(y);
        Transcript tr = make_tr(prefix, prefix_len, dst);
        return ipa_verify_point(key, load_g1(commitment), load_fr(point), pf, &tr) ? 1 : 0;
    } catch (const std::exception&) {
        return -1;
    }
}
int orc_ipa_prove_commitment(const uint8_t* bases, uint64_t N, const uint8_t* a, uint64_t len, const uint8_t* commitment,
                             uint8_t* L, uint8_t* R, uint8_t* tip) {
    ORC_TRY
    IpaKey key(load_g1s(bases, N + 1), N);
    LagrangeBasis data = LagrangeBasis::from_vec(load_frs(a, len));
    IpaCommitProof pf = ipa_prove_commitment(key, load_g1(commitment), data);
    store_g1s(L, pf.l);
    store_g1s(R, pf.r);
    store_fr(tip, pf.tip);
    ORC_CATCH
}
int orc_ipa_verify_commitment(const uint8_t* bases, uint64_t N, const uint8_t* commitment, const uint8_t* L, const uint8_t* R,
                              uint64_t rounds, const uint8_t* tip) {
    try {
        IpaKey key(load_g1s(bases, N + 1), N);
        IpaCommitProof pf;
        pf.l = load_g1s(L, rounds);
        pf.r = load_g1s(R, rounds);
        pf.tip = load_fr(tip);
        return ipa_verify_commitment_proof(key, load_g1(commitment), pf) ? 1 : 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// ------------------------------------------------------------------ IPA CRS (ipa_point_generator.rs:51-109)
int orc_ipa_crs_gen(const uint8_t* seed, uint64_t seed_len, uint64_t num, uint8_t* out /* num points */, uint64_t* next_index) {
    ORC_TRY
    std::vector<G1Affine> v = ipa_crs_gen(std::vector<uint8_t>(seed, seed + seed_len), num, next_index);
    for (size_t i = 0; i < v.size(); ++i) store_aff(out + 64 * i, v[i]);
    ORC_CATCH
}
// gen_at: 1 = a point (written to out), 0 = PointGeneratorError::InvalidPoint
int orc_ipa_crs_gen_at(const uint8_t* seed, uint64_t seed_len, uint64_t index, uint8_t* out) {
    try {
        uint8_t le[8];
        for (int k = 0; k < 8; ++k) le[k] = (uint8_t)(index >> (8 * k));
        G1Affine pt;
        if (!eth_hash_to_curve(std::vector<uint8_t>(seed, seed + seed_len), le, 8, pt)) return 0;
        store_aff(out, pt);
        return 1;
    } catch (const std::exception&) {
        return -1;
    }
}

// ------------------------------------------------------------------ KZG (K3; K4 restated with the known tau)
int orc_kzg_setup(uint64_t max_items, const uint8_t* tau, uint8_t* out_lagrange /* next_pow2(max_items) points */) {
    ORC_TRY
    KzgKey key = kzg_setup(max_items, load_fr(tau));
    store_g1s(out_lagrange, key.lagrange_commitments);
    ORC_CATCH
}
// ok_out: 1 = proof produced; 0 = the reference would panic (quirk Q2)
int orc_kzg_prove(const uint8_t* lagrange, uint64_t n, const uint8_t* data, uint64_t len, const uint8_t* point,
                  uint8_t* proof, uint8_t* y, int* ok_out) {
    ORC_TRY
    KzgKey key(load_g1s(lagrange, n), Fr::zero());
    LagrangeBasis lb = LagrangeBasis::from_vec(load_frs(data, len), n);
    bool ok;
    KzgProof pf = kzg_prove_point(key, load_fr(point), lb, &ok);
    *ok_out = ok ? 1 : 0;
    store_g1(proof, pf.proof);
    store_fr(y, pf.y);
    ORC_CATCH
}
int orc_kzg_verify_tau(const uint8_t* lagrange, uint64_t n, const uint8_t* tau, const uint8_t* commitment,
                       const uint8_t* point, const uint8_t* proof, const uint8_t* y) {
    try {
        KzgKey key(load_g1s(lagrange, n), load_fr(tau));
        KzgProof pf;
        pf.proof = load_g1(proof);
        pf.y = load_fr(y);
        return kzg_verify_point_with_tau(key, load_g1(commitment), load_fr(point), pf) ? 1 : 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// ------------------------------------------------------------------ multiproof (P1, P2)
// scheme 0 = IPA (bases: N+1 points), 1 = KZG (bases: N Lagrange points).  f[m][N], C[m], z[m], y[m].
// IPA outputs: D, L[log2 N], R[log2 N], tip, yout.  KZG outputs: D, L[0] = proof point, yout.
int orc_multiproof_prove(int scheme, const uint8_t* bases, uint64_t N, const uint8_t* f, const uint8_t* C, const uint64_t* z,
                         const uint8_t* y, uint64_t m, uint8_t* D, uint8_t* L, uint8_t* R, uint8_t* tip, uint8_t* yout) {
    ORC_TRY
    std::vector<LagrangeBasis> datas(m);
    std::vector<ProverQuery> qs(m);
    for (uint64_t k = 0; k < m; ++k) {
        datas[k] = LagrangeBasis::from_vec(load_frs(f + 32 * N * k, N));
        qs[k].data = &datas[k];
        qs[k].commit = load_g1(C + 64 * k);
        qs[k].z = z[k];
        qs[k].y = load_fr(y + 32 * k);
    }
    if (scheme == 0) {
        IpaKey key(load_g1s(bases, N + 1), N);
        IpaMultiproof mp = ipa_prove_multiproof(key, qs);
        store_g1(D, mp.d);
        store_g1s(L, mp.proof.l);
        store_g1s(R, mp.proof.r);
        store_fr(tip, mp.proof.tip);
        store_fr(yout, mp.proof.y);
    } else {
        KzgKey key(load_g1s(bases, N), Fr::zero());
        KzgMultiproof mp = kzg_prove_multiproof(key, qs);
        store_g1(D, mp.d);
        store_g1(L, mp.proof.proof);
        store_fr(yout, mp.proof.y);
    }
    ORC_CATCH
}
int orc_multiproof_verify(int scheme, const uint8_t* bases, uint64_t N, const uint8_t* tau, const uint8_t* C, const uint64_t* z,
                          const uint8_t* y, uint64_t m, const uint8_t* D, const uint8_t* L, const uint8_t* R, uint64_t rounds,
                          const uint8_t* tip, const uint8_t* yproof) {
    try {
        std::vector<VerifierQuery> qs(m);
        for (uint64_t k = 0; k < m; ++k) {
            qs[k].commit = load_g1(C + 64 * k);
            qs[k].z = z[k];
            qs[k].y = load_fr(y + 32 * k);
        }
        if (scheme == 0) {
            IpaKey key(load_g1s(bases, N + 1), N);
            IpaMultiproof mp;
            mp.d = load_g1(D);
            mp.proof.l = load_g1s(L, rounds);
            mp.proof.r = load_g1s(R, rounds);
            mp.proof.tip = load_fr(tip);
            mp.proof.y = load_fr(yproof);
            return ipa_verify_multiproof(key, qs, mp) ? 1 : 0;
        } else {
            KzgKey key(load_g1s(bases, N), load_fr(tau));
            KzgMultiproof mp;
            mp.d = load_g1(D);
            mp.proof.proof = load_g1(L);
            mp.proof.y = load_fr(yproof);
            return kzg_verify_multiproof_with_tau(key, qs, mp) ? 1 : 0;
        }
    } catch (const std::exception&) {
        return -1;
    }
}

// ------------------------------------------------------------------ verkle tree (T1)
int orc_tree_commit(const uint8_t* bases, uint64_t nb, const uint8_t* keys, uint64_t key_len, const uint8_t* values,
                    uint64_t n, uint64_t ext_width, uint8_t* out) {
    ORC_TRY
    std::vector<G1> b = load_g1s(bases, nb);
    Tree t(key_len);
    for (uint64_t i = 0; i < n; ++i) t.insert(keys + key_len * i, values + 32 * i);
    store_g1(out, t.commitment(b, ext_width));
    ORC_CATCH
}

}  // extern "C"
