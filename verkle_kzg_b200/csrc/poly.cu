// E1 / K1 / K2 / K3: evaluation-form polynomial work of the reference, batched over B openings, one
// warp per opening.
//   LagrangeBasis::evaluate                       lagrange_basis.rs:63-83   (3-way branch, quirk Q2)
//   LagrangeBasis::divide_by_vanishing            lagrange_basis.rs:91-119  (in-domain quotient)
//   LagrangeBasis::divive_by_vanishing_outside... lagrange_basis.rs:121-142
//   KZG::prove_point                              kzg/mod.rs:136-154        (quotient, then M1 over the SRS)
// The reference spends two field inversions per element in the in-domain quotient; every one of those
// denominators is a difference of two roots of unity, w^i - w^j = w^j (w^(i-j) - 1), so a key-load table
// of 1/(w^d - 1) replaces them (field arithmetic is exact: the products are the same field elements).
// Shapes: data rows hold `len` evaluations (LagrangeBasis::from_vec: domain size Dn = next_pow2(len));
// the key's precompute has size N with domain next_pow2(N).  The reference indexes out of bounds unless
// Dn <= N, so that is required here.
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

int32_t build_domain(vkzg_ctx* ctx, uint32_t log2n, uint32_t size_n, DomainTables& d);

__device__ __forceinline__ fp_t shfl_xor_fp2(const fp_t& v, int mask) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], mask);
    return r;
}
__device__ __forceinline__ fp_t warp_sum_fr2(fp_t v) {
#pragma unroll 1
    for (int m = 16; m > 0; m >>= 1) v = fp_add<S>(v, shfl_xor_fp2(v, m));
    return v;
}

struct PolyArgs {
    const fp_t* f;       // [B / share][len]
    const fp_t* points;  // [B], or nullptr: opening p is at the in-domain index p % share
    uint64_t B;
    uint32_t share;      // consecutive openings per data row (1: a row per opening; Dn: all in-domain openings of each row)
    uint32_t len, Dn;    // data row length, data domain size
    uint32_t N, Np;      // key size, key domain size
    const fp_t *wD, *wD_inv, *dD_inv;  // data-domain tables
    const fp_t* wK;                    // key-domain w^i
    fp_t n_inv;                        // 1 / N
    fp_t* q;             // [B][Dn] or nullptr (evaluate only)
    fp_t* scratch;       // [B][len] when q == nullptr
    fp_t* y;             // [B]
    int32_t* err;        // set to 1 if a row hits the reference's out-of-bounds panic
};

__global__ void __launch_bounds__(128) k_poly(PolyArgs A) {
    uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (p >= A.B) return;
    const fp_t* f = A.f + (A.share > 1 ? p / A.share : p) * A.len;
    fp_t* row = A.q ? A.q + p * A.Dn : A.scratch + p * A.len;
    const fp_t z = A.points ? fp_load_ro(A.points + p) : fp_from_u32<S>((uint32_t)(p % A.share));
    const fp_t zc = fp_from_mont<S>(z);
    const bool small = (zc.l[1] | zc.l[2] | zc.l[3] | zc.l[4] | zc.l[5] | zc.l[6] | zc.l[7]) == 0;
    const uint32_t zi = zc.l[0];

    // ---- evaluate (lagrange_basis.rs:63-72)
    fp_t y;
    if (small && zi <= A.len - 1) {
        y = fp_load_ro(f + zi);
    } else if (small && zi <= A.Dn) {
        y = fp_zero<S>();
    } else if (small && zi < A.N) {
        // barycentric coefficients are the unit vector e_zi (precompute.rs:74-78); zi > Dn >= len here
        y = fp_zero<S>();
    } else {
        // t * sum_i f_i w^i / (z - w^i) over the key's domain (precompute.rs:80-87), i < min(len, N) = len
        fp_t zp = fp_one<S>();
        for (int b = 31 - __clz(A.N); b >= 0; --b) {
            zp = fp_mul_ni<S>(zp, zp);
            if ((A.N >> b) & 1) zp = fp_mul_ni<S>(zp, z);
        }
        fp_t t = fp_mul_ni<S>(fp_sub<S>(zp, fp_one<S>()), A.n_inv);
        fp_t run = fp_one<S>();
        for (uint32_t i = lane; i < A.len; i += 32) {
            fp_store(row + i, run);
            run = fp_mul_ni<S>(run, fp_sub<S>(z, fp_load_ro(A.wK + i)));
        }
        fp_t inv = warp_inverse_of_lane_products(run);
        fp_t acc = fp_zero<S>();
        uint32_t cnt = A.len > lane ? (A.len - lane + 31) / 32 : 0;
        for (uint32_t k = cnt; k-- > 0;) {
            uint32_t i = lane + 32 * k;
            fp_t wi = fp_load_ro(A.wK + i);
            fp_t dinv = fp_mul_ni<S>(inv, fp_load(row + i));
            inv = fp_mul_ni<S>(inv, fp_sub<S>(z, wi));
            acc = fp_add<S>(acc, fp_mul_ni<S>(fp_mul_ni<S>(fp_load_ro(f + i), wi), dinv));
        }
        y = fp_mul_ni<S>(warp_sum_fr2(acc), t);
        __syncwarp();
    }
    if (lane == 0) fp_store(A.y + p, y);
    if (!A.q) return;

    if (small && zi <= A.N) {
        // ---- in-domain quotient (kzg/mod.rs:144-146 -> lagrange_basis.rs:91-119)
        const uint32_t idx = zi;
        if (idx >= A.N || idx >= A.Dn) {  // reference: index out of bounds panic (quirk Q2)
            for (uint32_t i = lane; i < A.Dn; i += 32) fp_store(row + i, fp_zero<S>());
            if (lane == 0) *A.err = 1;
            return;
        }
        const fp_t eval = idx < A.len ? fp_load_ro(f + idx) : fp_zero<S>();
        const fp_t w_inv_idx = fp_load_ro(A.wD_inv + idx);
        fp_t diag = fp_zero<S>();
        for (uint32_t i = lane; i < A.Dn; i += 32) {
            if (i == idx) continue;
            fp_t fi = i < A.len ? fp_load_ro(f + i) : fp_zero<S>();
            fp_t sub = fp_sub<S>(fi, eval);
            // 1 / (w^i - w^idx) = w^-idx / (w^(i-idx) - 1)
            fp_t qi = fp_mul_ni<S>(fp_mul_ni<S>(sub, w_inv_idx), fp_load_ro(A.dD_inv + ((i - idx) & (A.Dn - 1))));
            fp_store(row + i, qi);
            // A'(w^idx) / A'(w^i) = w_K^(i-idx);  1/(w^idx - w^i) = -1/(w^i - w^idx)
            diag = fp_sub<S>(diag, fp_mul_ni<S>(qi, fp_load_ro(A.wK + ((i - idx) & (A.Np - 1)))));
        }
        diag = warp_sum_fr2(diag);
        if (lane == 0) fp_store(row + idx, diag);
    } else {
        // ---- outside-domain quotient (lagrange_basis.rs:121-142); zero denominators stay zero like
        //      ark_ff::batch_inversion
        __syncwarp();
        fp_t run = fp_one<S>();
        for (uint32_t i = lane; i < A.Dn; i += 32) {
            fp_store(row + i, run);
            fp_t d = fp_sub<S>(fp_load_ro(A.wD + i), z);
            if (!fp_is_zero(d)) run = fp_mul_ni<S>(run, d);
        }
        fp_t inv = warp_inverse_of_lane_products(run);
        uint32_t cnt = A.Dn > lane ? (A.Dn - lane + 31) / 32 : 0;
        for (uint32_t k = cnt; k-- > 0;) {
            uint32_t i = lane + 32 * k;
            fp_t d = fp_sub<S>(fp_load_ro(A.wD + i), z);
            fp_t qi = fp_zero<S>();
            if (!fp_is_zero(d)) {
                fp_t dinv = fp_mul_ni<S>(inv, fp_load(row + i));
                inv = fp_mul_ni<S>(inv, d);
                fp_t fi = i < A.len ? fp_load_ro(f + i) : fp_zero<S>();
                qi = fp_mul_ni<S>(fp_sub<S>(fi, y), dinv);
            }
            fp_store(row + i, qi);
        }
    }
}

static int32_t data_domain(vkzg_ctx* ctx, const Key& k, uint32_t len, uint32_t domain_n, uint32_t& Dn, const DomainTables*& dt) {
    uint32_t lg = 0;
    uint32_t want = domain_n ? domain_n : len;
    if (want < len) return VKZG_ERR_ARG;
    while ((1u << lg) < want) ++lg;
    Dn = 1u << lg;
    if (Dn == k.domain_n) {
        dt = &k.dom;
        return VKZG_OK;
    }
    auto it = ctx->domains.find(lg);
    if (it == ctx->domains.end()) {
        DomainTables d;
        VK_TRY(build_domain(ctx, lg, Dn, d));
        it = ctx->domains.emplace(lg, d).first;
    }
    dt = &it->second;
    return VKZG_OK;
}

// q == nullptr: evaluate only.  Returns VKZG_ERR_RANGE (after synchronising) if check_err and a row panicked.
int32_t poly_batch(vkzg_ctx* ctx, const Key& k, const fp_t* d_f, uint32_t len, uint32_t domain_n, const fp_t* d_points, uint64_t B,
                   fp_t* d_q, fp_t* d_y, bool check_err, uint32_t share) {
    if (B == 0) return VKZG_OK;
    if (len == 0 || len > k.n) return VKZG_ERR_RANGE;
    uint32_t Dn;
    const DomainTables* dt;
    VK_TRY(data_domain(ctx, k, len, domain_n, Dn, dt));
    if (Dn > k.n) return VKZG_ERR_UNSUPPORTED;
    DevBuf<fp_t> scratch;
    DevBuf<int32_t> err;
    if (!d_q) VK_TRY(scratch.alloc(ctx, B * len));
    VK_TRY(err.alloc(ctx, 1));
    VK_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int32_t), ctx->stream));
    PolyArgs A;
    A.f = d_f;
    A.points = d_points;
    A.B = B;
    A.share = share ? share : 1;
    A.len = len;
    A.Dn = Dn;
    A.N = k.n;
    A.Np = k.domain_n;
    A.wD = dt->omega;
    A.wD_inv = dt->omega_inv;
    A.dD_inv = dt->diff_inv;
    A.wK = k.dom.omega;
    A.n_inv = k.dom.n_inv;
    A.q = d_q;
    A.scratch = scratch.p;
    A.y = d_y;
    A.err = err.p;
    k_poly<<<ceil_div_u64(B * 32, 128), 128, 0, ctx->stream>>>(A);
    VK_TRY(launch_check(ctx));
    if (check_err) {
        int32_t h = 0;
        VK_CUDA(cudaMemcpyAsync(&h, err.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        VK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h) return VKZG_ERR_RANGE;
    }
    return VKZG_OK;
}

uint32_t data_domain_size(uint32_t len, uint32_t domain_n) {
    uint32_t want = domain_n > len ? domain_n : len;
    uint32_t d = 1;
    while (d < want) d <<= 1;
    return d;
}

int32_t kzg_open_core(vkzg_ctx* ctx, const Key& k, const fp_t* d_f, uint32_t len, uint32_t domain_n, const fp_t* d_points,
                      uint64_t B, affine_t* d_proof, fp_t* d_y, bool check_err) {
    if (B == 0) return VKZG_OK;
    uint32_t Dn = data_domain_size(len, domain_n);
    if (Dn > k.n) return VKZG_ERR_UNSUPPORTED;
    DevBuf<fp_t> q;
    DevBuf<xyzz_t> acc;
    VK_TRY(q.alloc(ctx, B * Dn));
    VK_TRY(acc.alloc(ctx, B));
    VK_TRY(poly_batch(ctx, k, d_f, len, domain_n, d_points, B, q, d_y, check_err));
    VK_TRY(fixed_base_msm(ctx, k, q, Dn, B, 0, 0xffffffffu, acc));  // inner_product zips Dn <= N terms
    return normalize_points(ctx, acc, B, d_proof);
}

// L1 / I2: element-wise Fr vector arithmetic (lagrange_basis.rs:202-233 AddAssign / Sub / Mul<F>, utils.rs:21-38
// elementwise_mul / vec_add_and_distribute).  Pure streaming: 64-96 bytes moved per 0-1 multiplication, HBM-bound.
__global__ void __launch_bounds__(256) k_fr_vec(int op, const fp_t* __restrict__ a, const fp_t* __restrict__ b, fp_t x, uint64_t n,
                                                fp_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t va = fp_load_ro(a + i), r;
    switch (op) {
        case 0: r = fp_add<S>(va, fp_load_ro(b + i)); break;                       // a + b
        case 1: r = fp_sub<S>(va, fp_load_ro(b + i)); break;                       // a - b
        case 2: r = fp_mul<S>(va, fp_load_ro(b + i)); break;                       // a .* b
        case 3: r = fp_mul<S>(va, x); break;                                       // a * x
        default: r = fp_add<S>(va, fp_mul<S>(x, fp_load_ro(b + i))); break;        // a + x * b
    }
    fp_store(out + i, r);
}

}  // namespace vk

using namespace vk;

extern "C" {

int32_t vkzg_fr_vector_op_dev(vkzg_ctx* ctx, int32_t op, const vkzg_fr* d_a, const vkzg_fr* d_b, const vkzg_fr* x, uint64_t n,
                              vkzg_fr* d_out) {
    VK_TRY(ctx_check(ctx));
    if (op < 0 || op > 4 || (n && (!d_a || !d_out))) return VKZG_ERR_ARG;
    if (n && op != 3 && !d_b) return VKZG_ERR_ARG;
    if ((op == 3 || op == 4) && !x) return VKZG_ERR_ARG;
    if (!n) return VKZG_OK;
    fp_t xv = fp_zero<S>();
    if (x) memcpy(&xv, x, sizeof(xv));
    k_fr_vec<<<ceil_div_u64(n, 256), 256, 0, ctx->stream>>>(op, (const fp_t*)d_a, (const fp_t*)(d_b ? d_b : d_a), xv, n, (fp_t*)d_out);
    return launch_check(ctx);
}

int32_t vkzg_fr_vector_op(vkzg_ctx* ctx, int32_t op, const vkzg_fr* a, const vkzg_fr* b, const vkzg_fr* x, uint64_t n, vkzg_fr* out) {
    VK_TRY(ctx_check(ctx));
    if (n && (!a || !out)) return VKZG_ERR_ARG;
    DevBuf<fp_t> da, db, dout;
    VK_TRY(upload(ctx, da, a, n));
    if (b) VK_TRY(upload(ctx, db, b, n));
    VK_TRY(dout.alloc(ctx, n));
    VK_TRY(vkzg_fr_vector_op_dev(ctx, op, (const vkzg_fr*)da.p, b ? (const vkzg_fr*)db.p : nullptr, x, n, (vkzg_fr*)dout.p));
    VK_TRY(download(ctx, out, dout.p, n));
    return stream_sync(ctx);
}

int32_t vkzg_evaluate_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n, const vkzg_fr* points,
                            uint64_t B, vkzg_fr* out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!f || !points || !out))) return VKZG_ERR_ARG;
    DevBuf<fp_t> df, dp, dy;
    VK_TRY(upload(ctx, df, f, B * len));
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(dy.alloc(ctx, B));
    VK_TRY(poly_batch(ctx, *k, df, len, domain_n, dp, B, nullptr, dy, false));
    VK_TRY(download(ctx, out, dy.p, B));
    return stream_sync(ctx);
}

int32_t vkzg_quotient_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n, const vkzg_fr* points,
                            uint64_t B, vkzg_fr* out, vkzg_fr* y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!f || !points || !out || !y))) return VKZG_ERR_ARG;
    if (len == 0 || len > k->n) return VKZG_ERR_RANGE;
    uint32_t Dn = data_domain_size(len, domain_n);
    DevBuf<fp_t> df, dp, dq, dy;
    VK_TRY(upload(ctx, df, f, B * len));
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(dq.alloc(ctx, B * Dn));
    VK_TRY(dy.alloc(ctx, B));
    VK_TRY(poly_batch(ctx, *k, df, len, domain_n, dp, B, dq, dy, true));
    VK_TRY(download(ctx, out, dq.p, B * Dn));
    VK_TRY(download(ctx, y, dy.p, B));
    return stream_sync(ctx);
}

// device-pointer variants (rows that hit the reference's panic get a zero quotient; no status)
int32_t vkzg_quotient_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_f, uint32_t len, uint32_t domain_n,
                                const vkzg_fr* d_points, uint64_t B, vkzg_fr* d_out, vkzg_fr* d_y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!d_f || !d_points || !d_out || !d_y))) return VKZG_ERR_ARG;
    return poly_batch(ctx, *k, (const fp_t*)d_f, len, domain_n, (const fp_t*)d_points, B, (fp_t*)d_out, (fp_t*)d_y, false);
}
int32_t vkzg_evaluate_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_f, uint32_t len, uint32_t domain_n,
                                const vkzg_fr* d_points, uint64_t B, vkzg_fr* d_y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!d_f || !d_points || !d_y))) return VKZG_ERR_ARG;
    return poly_batch(ctx, *k, (const fp_t*)d_f, len, domain_n, (const fp_t*)d_points, B, nullptr, (fp_t*)d_y, false);
}

int32_t vkzg_kzg_open_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_f, uint32_t len, uint32_t domain_n,
                                const vkzg_fr* d_points, uint64_t B, vkzg_g1_affine* d_proof, vkzg_fr* d_y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!d_f || !d_points || !d_proof || !d_y))) return VKZG_ERR_ARG;
    if (len == 0 || len > k->n) return VKZG_ERR_RANGE;
    return kzg_open_core(ctx, *k, (const fp_t*)d_f, len, domain_n, (const fp_t*)d_points, B, (affine_t*)d_proof, (fp_t*)d_y, false);
}

int32_t vkzg_kzg_open_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n, const vkzg_fr* points,
                            uint64_t B, vkzg_g1_affine* proof, vkzg_fr* y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!f || !points || !proof || !y))) return VKZG_ERR_ARG;
    if (len == 0 || len > k->n) return VKZG_ERR_RANGE;
    DevBuf<fp_t> df, dp, dy;
    DevBuf<affine_t> dpr;
    VK_TRY(upload(ctx, df, f, B * len));
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(dpr.alloc(ctx, B));
    VK_TRY(dy.alloc(ctx, B));
    VK_TRY(kzg_open_core(ctx, *k, df, len, domain_n, dp, B, dpr, dy, true));
    VK_TRY(download(ctx, proof, dpr.p, B));
    VK_TRY(download(ctx, y, dy.p, B));
    return stream_sync(ctx);
}

// KZG::prove_all_points (kzg/mod.rs:200-235): the openings of every vector at ALL points of its data domain.  The reference's
// function is unreachable (private, its test is not registered), stops after the h-vector of Feist-Khovratovich (no final FFT)
// and indexes data[i] for i up to 2N - 1; what is reproduced is its contract — proof[b][i] == prove_point(f_b, i), y[b][i] = f_b[i].
// With the window tables an opening is Dn x W table additions: Dn openings of a width-256 row cost the same field products as
// FK's ~3800 variable-base scalar multiplications (two size-2Dn group FFTs + the Hadamard product), so the all-points prover is
// the batched single-point path with the data row shared by Dn consecutive openings, cut into pieces of <= 1 GiB of quotients.
int32_t vkzg_kzg_prove_all_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n, uint64_t B,
                                 vkzg_g1_affine* proof, vkzg_fr* y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!f || !proof || !y))) return VKZG_ERR_ARG;
    if (len == 0 || len > k->n) return VKZG_ERR_RANGE;
    if (domain_n && domain_n < len) return VKZG_ERR_ARG;
    const uint32_t Dn = data_domain_size(len, domain_n);
    if (Dn > k->n) return VKZG_ERR_UNSUPPORTED;
    if (B == 0) return VKZG_OK;
    uint64_t piece = (1ull << 30) / ((uint64_t)Dn * Dn * sizeof(fp_t));
    if (piece == 0) piece = 1;
    if (piece > B) piece = B;
    DevBuf<fp_t> df, dq, dy;
    DevBuf<xyzz_t> acc;
    DevBuf<affine_t> dpr;
    VK_TRY(upload(ctx, df, f, B * len));
    VK_TRY(dq.alloc(ctx, piece * Dn * Dn));
    VK_TRY(dy.alloc(ctx, B * Dn));
    VK_TRY(acc.alloc(ctx, piece * Dn));
    VK_TRY(dpr.alloc(ctx, B * Dn));
    for (uint64_t b0 = 0; b0 < B; b0 += piece) {
        const uint64_t nb = B - b0 < piece ? B - b0 : piece;
        VK_TRY(poly_batch(ctx, *k, df.p + b0 * len, len, domain_n, nullptr, nb * Dn, dq, dy.p + b0 * Dn, false, Dn));
        VK_TRY(fixed_base_msm(ctx, *k, dq, Dn, nb * Dn, 0, 0xffffffffu, acc));
        VK_TRY(normalize_points(ctx, acc, nb * Dn, dpr.p + b0 * Dn));
    }
    VK_TRY(download(ctx, proof, dpr.p, B * Dn));
    VK_TRY(download(ctx, y, dy.p, B * Dn));
    return stream_sync(ctx);
}

// KZG::commit + KZG::prove_point for bulk callers that commit to AND open the same vectors (the reference's own bench shape,
// benches/kzg.rs): the rows cross PCIe once, chunk by chunk under the commitment MSM of the previous chunk.
int32_t vkzg_kzg_commit_open_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n,
                                   const vkzg_fr* points, uint64_t B, vkzg_g1_affine* commitments, vkzg_g1_affine* proof, vkzg_fr* y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!f || !points || !commitments || !proof || !y))) return VKZG_ERR_ARG;
    if (len == 0 || len > k->n) return VKZG_ERR_RANGE;
    if (B == 0) return VKZG_OK;
    DevBuf<fp_t> df, dp, dy;
    DevBuf<affine_t> dc, dpr;
    DevBuf<xyzz_t> acc;
    VK_TRY(df.alloc(ctx, B * len));
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(dc.alloc(ctx, B));
    VK_TRY(dpr.alloc(ctx, B));
    VK_TRY(dy.alloc(ctx, B));
    VK_TRY(acc.alloc(ctx, B));
    ChunkedUpload up(ctx);
    VK_TRY(up.init());
    for (uint64_t b0 = 0, nb = 0; b0 < B; b0 += nb) {
        nb = pipeline_piece(B, b0);
        VK_TRY(up.copy(df.p + b0 * len, (const fp_t*)f + b0 * len, nb * len * sizeof(fp_t)));
        VK_TRY(up.publish());
        VK_TRY(fixed_base_msm(ctx, *k, df.p + b0 * len, len, nb, 0, 0xffffffffu, acc.p + b0));
    }
    VK_TRY(normalize_points(ctx, acc, B, dc));
    VK_TRY(kzg_open_core(ctx, *k, df, len, domain_n, dp, B, dpr, dy, true));
    VK_TRY(download(ctx, commitments, dc.p, B));
    VK_TRY(download(ctx, proof, dpr.p, B));
    VK_TRY(download(ctx, y, dy.p, B));
    return stream_sync(ctx);
}

}  // extern "C"
