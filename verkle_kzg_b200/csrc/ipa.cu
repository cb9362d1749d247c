// I1 / I3 / I4: the IPA opening of the reference (vector-commit/src/ipa/mod.rs), batched over B
// independent proofs.
//
// low_level_ipa (ipa/mod.rs:268-319) halves a, b and the generator vector G every round, which costs m
// variable-base scalar multiplications by the round challenge to fold G.  Group arithmetic is exact, so
// the folded generators never have to exist: after k rounds
//     g^(k)_j = sum over the original bases i with  i mod len_k == j  of  coef_k(i) * G_i,
//     coef_{k+1}(i) = coef_k(i) * (x_k  if  i mod len_k < m  else 1)          (g' = g_R + x g_L, :309)
// and both cross terms of a round are fixed-base MSMs over the ORIGINAL key, half of the bases each:
//     L_k = sum_{i mod len_k >= m} (coef_k(i) a[i mod m]) G_i      + (w <a_L, b_R>) Q         (:299)
//     R_k = sum_{i mod len_k <  m} (coef_k(i) a[m + i mod m]) G_i  + (w <a_R, b_L>) Q         (:300)
// (q <- q * w at :294 turns into the scalar factor w because Q is fixed too).  One round is therefore
//   k_ipa_fold_scalars   warp per proof: fold a, b, coef by the previous challenge, emit the 2 x (N/2 + 1)
//                        MSM scalars of this round
//   k_fixed_base_msm     (commit.cu) 2 warps per proof: L and R through the key's window tables
//   k_ipa_challenge      thread per proof: L, R to affine with one shared inversion, append to the
//                        transcript, squeeze x_k (SHA-256 XMD on the device)
// i.e. the whole round loop runs on the device with no host round trip; outputs are canonical affine.
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

__device__ __forceinline__ fp_t shfl_xor_fp(const fp_t& v, int mask) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], mask);
    return r;
}
__device__ __forceinline__ fp_t warp_sum_fr(fp_t v) {
#pragma unroll 1
    for (int m = 16; m > 0; m >>= 1) v = fp_add<S>(v, shfl_xor_fp(v, m));
    return v;
}

// canonical(a) < n  (n small)
__device__ __forceinline__ bool canon_lt_u32(const fp_t& canon, uint32_t n) {
    return (canon.l[1] | canon.l[2] | canon.l[3] | canon.l[4] | canon.l[5] | canon.l[6] | canon.l[7]) == 0 && canon.l[0] < n;
}

// ---------------------------------------------------------------------------------------------------
// B1: PrecomputedLagrange::compute_barycentric_coefficients (precompute.rs:72-90).  Warp per point.
//   point < size (strict, quirk Q3): unit vector;  else b_i = ((z^size - 1)/size) w^i / (z - w^i)
// The N inversions are ONE inversion per warp: lanes own i = lane, lane+32, ... and batch them, the lane products are
// inverted together (warp_inverse_of_lane_products).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_barycentric(const fp_t* __restrict__ points, uint64_t B, uint32_t size,
                                                     const fp_t* __restrict__ omega, fp_t n_inv, fp_t* __restrict__ out) {
    uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (p >= B) return;
    fp_t z = fp_load_ro(points + p);
    fp_t zc = fp_from_mont<S>(z);
    fp_t* o = out + p * size;
    if (canon_lt_u32(zc, size)) {
        for (uint32_t i = lane; i < size; i += 32) fp_store(o + i, i == zc.l[0] ? fp_one<S>() : fp_zero<S>());
        return;
    }
    // t = (z^size - 1) / size
    fp_t zp = fp_one<S>();
    for (int b = 31 - __clz(size); b >= 0; --b) {
        zp = fp_mul_ni<S>(zp, zp);
        if ((size >> b) & 1) zp = fp_mul_ni<S>(zp, z);
    }
    fp_t t = fp_mul_ni<S>(fp_sub<S>(zp, fp_one<S>()), n_inv);
    // per-lane batch inversion of (z - w^i), prefix products parked in the output row
    fp_t run = fp_one<S>();
    for (uint32_t i = lane; i < size; i += 32) {
        fp_store(o + i, run);
        run = fp_mul_ni<S>(run, fp_sub<S>(z, fp_load_ro(omega + i)));
    }
    fp_t inv = warp_inverse_of_lane_products(run);
    uint32_t cnt = size > lane ? (size - lane + 31) / 32 : 0;
    for (uint32_t k = cnt; k-- > 0;) {
        uint32_t i = lane + 32 * k;
        fp_t wi = fp_load_ro(omega + i);
        fp_t dinv = fp_mul_ni<S>(inv, fp_load(o + i));
        inv = fp_mul_ni<S>(inv, fp_sub<S>(z, wi));
        fp_store(o + i, fp_mul_ni<S>(fp_mul_ni<S>(t, wi), dinv));
    }
}

int32_t barycentric_batch(vkzg_ctx* ctx, const Key& k, const fp_t* d_points, uint64_t B, fp_t* d_out) {
    if (!B) return VKZG_OK;
    k_barycentric<<<ceil_div_u64(B * 32, 128), 128, 0, ctx->stream>>>(d_points, B, k.n, k.dom.omega, k.dom.n_inv, d_out);
    return launch_check(ctx);
}

// ---------------------------------------------------------------------------------------------------
// per-proof state
// ---------------------------------------------------------------------------------------------------
struct IpaState {
    fp_t* a;        // [B][N]  folded in place
    fp_t* b;        // [B][N]
    fp_t* coef;     // [B][N]
    fp_t* sc;       // [2B][T] MSM scalars of the current round
    fp_t* x;        // [B] challenge of the previous round
    fp_t* w;        // [B]
    fp_t* y;        // [B] evaluation
    transcript_t* tr;  // [B]
    xyzz_t* lr;     // [2B]
};

// y = <a, b>  (ipa/mod.rs:277), coef = 1.  Warp per proof.
__global__ void __launch_bounds__(128) k_ipa_begin(IpaState st, uint64_t B, uint32_t N, bool with_b, fp_t* __restrict__ y_out) {
    uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (p >= B) return;
    fp_t acc = fp_zero<S>();
    for (uint32_t i = lane; i < N; i += 32) {
        fp_store(st.coef + p * N + i, fp_one<S>());
        if (with_b) acc = fp_add<S>(acc, fp_mul<S>(fp_load(st.a + p * N + i), fp_load(st.b + p * N + i)));
    }
    if (with_b) {
        acc = warp_sum_fr(acc);
        if (lane == 0) {
            fp_store(st.y + p, acc);
            fp_store(y_out + p, acc);
        }
    }
}

struct TrPrefix {
    uint8_t bytes[TR_PREFIX_INLINE];
    uint32_t len;
    uint8_t dst[TR_DST_MAX];
    uint32_t dst_len;
    uint32_t mid[8];      // SHA-256 state after the first mid_bytes bytes of Z_pad || prefix (long prefixes, pre-hashed on the host)
    uint32_t mid_bytes;
};
__host__ __device__ inline void tr_begin(transcript_t& t, const TrPrefix& pre) {
    t.len = 0;
    tr_append_raw(t, pre.bytes, pre.len);
    t.dst_len = pre.dst_len;
    for (uint32_t i = 0; i < pre.dst_len; ++i) t.dst[i] = pre.dst[i];
    for (int i = 0; i < 8; ++i) t.mid[i] = pre.mid[i];
    t.mid_bytes = pre.mid_bytes;
}

// transcript start (ipa/mod.rs:286-292): append C, input point, output point; w = digest("w").
// mode 1 = prove_commitment (ipa/mod.rs:210-213): append C; digest("x") (result unused).
// prefix_each (may be null): a different in-flight transcript state per proof, each_len <= TR_PREFIX_INLINE bytes each
// (batched multiproofs: every inner opening continues its own outer transcript)
__global__ void __launch_bounds__(64) k_ipa_transcript_begin(IpaState st, uint64_t B, const affine_t* __restrict__ C,
                                                             const fp_t* __restrict__ points, TrPrefix pre, int mode,
                                                             const uint8_t* __restrict__ prefix_each, uint32_t each_len) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    transcript_t t;
    tr_begin(t, pre);
    if (prefix_each) tr_append_raw(t, prefix_each + p * each_len, each_len);
    affine_t c;
    c.x = fp_load(&C[p].x);
    c.y = fp_load(&C[p].y);
    tr_append_point(t, c, "C");
    if (mode == 0) {
        tr_append_fr(t, fp_load(points + p), "input point");
        tr_append_fr(t, fp_load(st.y + p), "output point");
        fp_store(st.w + p, tr_digest(t, "w"));
    } else {
        tr_digest(t, "x");
        fp_store(st.w + p, fp_zero<S>());
    }
    st.tr[p] = t;
}

// fold by the previous round's challenge (ipa/mod.rs:308-310), then the scalars of round `round`.
//   fold_m  = half length of the round just finished (0: nothing to fold)
//   m       = half length of this round (0: no scalars, final call -> tip)
// G threads per proof: 32 (a warp; big batches) or 256 (a whole CTA; a handful of proofs, where the kernel is a chain of
// dependent products per thread and eight times more threads make it eight times shorter).
template <int G>
__global__ void __launch_bounds__(G == 32 ? 128 : G) k_ipa_fold_scalars(IpaState st, uint64_t B, uint32_t N, uint32_t fold_m, uint32_t m,
                                                                        uint32_t T, bool with_b, fp_t* __restrict__ tip_out) {
    uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    uint32_t lane = threadIdx.x % G;
    if (p >= B) return;  // (G == 256: the whole CTA leaves together)
    fp_t* a = st.a + p * N;
    fp_t* b = st.b + p * N;
    fp_t* coef = st.coef + p * N;
    if (fold_m) {
        fp_t x = fp_load(st.x + p);
        for (uint32_t j = lane; j < fold_m; j += G) {
            fp_store(a + j, fp_add<S>(fp_load(a + j), fp_mul<S>(x, fp_load(a + fold_m + j))));
            if (with_b) fp_store(b + j, fp_add<S>(fp_load(b + fold_m + j), fp_mul<S>(x, fp_load(b + j))));
        }
        for (uint32_t i = lane; i < N; i += G)
            if ((i & (2 * fold_m - 1)) < fold_m) fp_store(coef + i, fp_mul<S>(fp_load(coef + i), x));
        if (G == 32)
            __syncwarp();
        else
            __syncthreads();
    }
    if (m == 0) {
        if (lane == 0) fp_store(tip_out + p, fp_load(a));
        return;
    }
    fp_t* scL = st.sc + (2 * p) * T;
    fp_t* scR = scL + T;
    const uint32_t half = N / 2;
    for (uint32_t j = lane; j < half; j += G) {
        uint32_t blk = (j / m) * 2 * m, r = j % m;
        fp_store(scL + j, fp_mul<S>(fp_load(coef + blk + m + r), fp_load(a + r)));
        fp_store(scR + j, fp_mul<S>(fp_load(coef + blk + r), fp_load(a + m + r)));
    }
    if (with_b) {
        fp_t ipl = fp_zero<S>(), ipr = fp_zero<S>();
        for (uint32_t j = lane; j < m; j += G) {
            ipl = fp_add<S>(ipl, fp_mul<S>(fp_load(a + j), fp_load(b + m + j)));
            ipr = fp_add<S>(ipr, fp_mul<S>(fp_load(a + m + j), fp_load(b + j)));
        }
        ipl = warp_sum_fr(ipl);
        ipr = warp_sum_fr(ipr);
        if (G > 32) {  // across the warps of the CTA
            __shared__ fp_t red[2][G / 32];
            if ((lane & 31) == 0) {
                red[0][lane >> 5] = ipl;
                red[1][lane >> 5] = ipr;
            }
            __syncthreads();
            if (lane == 0)
                for (int k = 1; k < G / 32; ++k) {
                    ipl = fp_add<S>(ipl, red[0][k]);
                    ipr = fp_add<S>(ipr, red[1][k]);
                }
        }
        if (lane == 0) {
            fp_t w = fp_load(st.w + p);
            fp_store(scL + half, fp_mul<S>(w, ipl));
            fp_store(scR + half, fp_mul<S>(w, ipr));
        }
    }
}

static int32_t launch_fold_scalars(vkzg_ctx* ctx, IpaState st, uint64_t B, uint32_t N, uint32_t fold_m, uint32_t m, uint32_t T,
                                   bool with_b, fp_t* tip_out) {
    if (B <= 64 && N >= 64)  // latency-bound: a CTA per proof
        k_ipa_fold_scalars<256><<<(uint32_t)B, 256, 0, ctx->stream>>>(st, B, N, fold_m, m, T, with_b, tip_out);
    else
        k_ipa_fold_scalars<32><<<ceil_div_u64(B * 32, 128), 128, 0, ctx->stream>>>(st, B, N, fold_m, m, T, with_b, tip_out);
    return launch_check(ctx);
}

// L, R -> affine (one inversion), outputs, transcript, challenge (ipa/mod.rs:301-306)
// SOLO (a handful of proofs): one thread per CTA, its own inversion — the shuffle scans of the shared inversion are ten
// products on the critical path that buy nothing when the other lanes are idle.
template <bool SOLO>
__global__ void __launch_bounds__(64) k_ipa_challenge(IpaState st, uint64_t B, uint32_t round, uint32_t rounds,
                                                      affine_t* __restrict__ L_out, affine_t* __restrict__ R_out) {
    uint64_t p = SOLO ? (uint64_t)blockIdx.x : (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < B && (!SOLO || threadIdx.x == 0);  // (no early return before the warp-wide inversion)
    if (SOLO && !live) return;
    xyzz_t l = xyzz_inf(), r = xyzz_inf();
    if (live) {
        l = st.lr[2 * p];
        r = st.lr[2 * p + 1];
    }
    bool linf = xyzz_is_inf(l), rinf = xyzz_is_inf(r);
    fp_t zl = linf ? fp_one<Q>() : l.zzz, zr = rinf ? fp_one<Q>() : r.zzz;
    // (batches: the 64 points of a warp share one inversion)
    fp_t inv = SOLO ? fp_inv<Q>(fp_mul_ni<Q>(zl, zr)) : warp_inverse_of_lane_products_t<Q>(fp_mul_ni<Q>(zl, zr));
    if (!live) return;
    affine_t la = linf ? affine_inf() : xyzz_to_affine_with_inv(l, fp_mul_ni<Q>(inv, zr));
    affine_t ra = rinf ? affine_inf() : xyzz_to_affine_with_inv(r, fp_mul_ni<Q>(inv, zl));
    L_out[p * rounds + round] = la;
    R_out[p * rounds + round] = ra;
    transcript_t t = st.tr[p];
    tr_append_point(t, la, "L");
    tr_append_point(t, ra, "R");
    fp_store(st.x + p, tr_digest(t, "x"));
    st.tr[p] = t;
}

static int32_t make_prefix(TrPrefix& pre, const uint8_t* prefix, uint32_t prefix_len, const char* dst) {
    memset(&pre, 0, sizeof(pre));
    if (prefix_len && !prefix) return VKZG_ERR_ARG;
    if (prefix_len > sizeof(pre.bytes)) {
        // the reference's transcript has no size limit (transcript.rs:34-52): absorb the whole 64-byte blocks of
        // Z_pad || prefix here, the device continues from that SHA-256 state with the < 64 bytes that are left
        const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        for (int i = 0; i < 8; ++i) pre.mid[i] = IV[i];
        const uint64_t total = (uint64_t)ARK04_Z_PAD_LEN + prefix_len;
        const uint64_t nblk = total / 64;
        uint8_t blk[64];
        for (uint64_t b = 0; b < nblk; ++b) {
            for (uint32_t i = 0; i < 64; ++i) {
                uint64_t pos = b * 64 + i;
                blk[i] = pos < ARK04_Z_PAD_LEN ? 0 : prefix[pos - ARK04_Z_PAD_LEN];
            }
            sha256_compress(pre.mid, blk);
        }
        pre.mid_bytes = (uint32_t)(nblk * 64);
        const uint32_t done = pre.mid_bytes - ARK04_Z_PAD_LEN;  // prefix bytes already absorbed
        pre.len = prefix_len - done;
        memcpy(pre.bytes, prefix + done, pre.len);
    } else {
        if (prefix_len) memcpy(pre.bytes, prefix, prefix_len);
        pre.len = prefix_len;
    }
    const char* d = dst ? dst : "ipa";
    size_t dl = strlen(d);
    if (dl == 0 || dl > TR_DST_MAX) return VKZG_ERR_ARG;
    memcpy(pre.dst, d, dl);
    pre.dst_len = (uint32_t)dl;
    return VKZG_OK;
}

static int32_t ipa_prove_one_stream(vkzg_ctx* ctx, const Key& k, int mode, uint32_t N, const fp_t* d_a, const fp_t* d_points,
                                    const affine_t* d_C, uint64_t B, const uint8_t* prefix, uint32_t prefix_len, const char* dst,
                                    affine_t* d_L, affine_t* d_R, fp_t* d_tip, fp_t* d_y, const uint8_t* d_prefix_each, uint32_t each_len);

// Big batches are proven as two half-batches on two streams: the per-round challenge / fold kernels are latency-bound
// (a thread or a warp per proof, 8 dependent rounds), so one half's hash-and-fold runs under the other half's MSM kernel.
int32_t ipa_prove_core(vkzg_ctx* ctx, const Key& k, int mode, uint32_t N, const fp_t* d_a, const fp_t* d_points,
                       const affine_t* d_C, uint64_t B, const uint8_t* prefix, uint32_t prefix_len, const char* dst, affine_t* d_L,
                       affine_t* d_R, fp_t* d_tip, fp_t* d_y, const uint8_t* d_prefix_each, uint32_t each_len) {
    if (d_prefix_each && (prefix_len || each_len > TR_PREFIX_INLINE)) return VKZG_ERR_ARG;
    if (!ctx->ipa_two_streams || B < 8192)
        return ipa_prove_one_stream(ctx, k, mode, N, d_a, d_points, d_C, B, prefix, prefix_len, dst, d_L, d_R, d_tip, d_y, d_prefix_each,
                                    each_len);
    if (!ctx->aux_stream) VK_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    if (!ctx->ev_fork) VK_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    if (!ctx->ev_join) VK_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    uint32_t rounds = 0;
    while ((1u << rounds) < N) ++rounds;
    const uint64_t B0 = B / 2, B1 = B - B0;
    cudaEvent_t fork = ctx->ev_fork, join = ctx->ev_join;
    VK_CUDA(cudaEventRecord(fork, ctx->stream));          // inputs are ready in main-stream order
    VK_CUDA(cudaStreamWaitEvent(ctx->aux_stream, fork, 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->aux_stream;                        // everything the second half enqueues (scratch included) goes to the aux stream
    int32_t st1 = ipa_prove_one_stream(ctx, k, mode, N, d_a + B0 * N, d_points ? d_points + B0 : nullptr, d_C + B0, B1, prefix, prefix_len,
                                       dst, d_L + B0 * rounds, d_R + B0 * rounds, d_tip + B0, d_y + B0,
                                       d_prefix_each ? d_prefix_each + B0 * each_len : nullptr, each_len);
    cudaEventRecord(join, ctx->stream);
    ctx->stream = main_stream;
    int32_t st0 = ipa_prove_one_stream(ctx, k, mode, N, d_a, d_points, d_C, B0, prefix, prefix_len, dst, d_L, d_R, d_tip, d_y, d_prefix_each,
                                       each_len);
    cudaStreamWaitEvent(ctx->stream, join, 0);
    return st0 != VKZG_OK ? st0 : st1;
}

// mode 0: prove_point (b from the evaluation point, q term); mode 1: prove_commitment (no b, no q).
// N = vector length (power of two, <= key size).
static int32_t ipa_prove_one_stream(vkzg_ctx* ctx, const Key& k, int mode, uint32_t N, const fp_t* d_a, const fp_t* d_points,
                       const affine_t* d_C, uint64_t B, const uint8_t* prefix, uint32_t prefix_len, const char* dst, affine_t* d_L,
                       affine_t* d_R, fp_t* d_tip, fp_t* d_y, const uint8_t* d_prefix_each, uint32_t each_len) {
    if (B == 0) return VKZG_OK;
    if (N < 2 || (N & (N - 1)) || N > k.n) return VKZG_ERR_UNSUPPORTED;
    const bool with_b = mode == 0;
    if (with_b && (!k.has_q || N != k.n)) return VKZG_ERR_UNSUPPORTED;
    TrPrefix pre;
    VK_TRY(make_prefix(pre, prefix, prefix_len, dst));
    uint32_t rounds = 0;
    while ((1u << rounds) < N) ++rounds;
    const uint32_t T = N / 2 + (with_b ? 1 : 0);
    cudaStream_t s = ctx->stream;
    DevBuf<fp_t> a, b, coef, sc, x, w, y;
    DevBuf<transcript_t> tr;
    DevBuf<xyzz_t> lr;
    VK_TRY(a.alloc(ctx, B * N));
    VK_TRY(b.alloc(ctx, with_b ? B * N : 1));
    VK_TRY(coef.alloc(ctx, B * N));
    VK_TRY(sc.alloc(ctx, 2 * B * T));
    VK_TRY(x.alloc(ctx, B));
    VK_TRY(w.alloc(ctx, B));
    VK_TRY(y.alloc(ctx, B));
    VK_TRY(tr.alloc(ctx, B));
    VK_TRY(lr.alloc(ctx, 2 * B));
    VK_CUDA(cudaMemcpyAsync(a.p, d_a, B * N * sizeof(fp_t), cudaMemcpyDeviceToDevice, s));
    IpaState st{a, b, coef, sc, x, w, y, tr, lr};
    if (with_b) VK_TRY(barycentric_batch(ctx, k, d_points, B, b));
    uint32_t wblocks = ceil_div_u64(B * 32, 128);
    k_ipa_begin<<<wblocks, 128, 0, s>>>(st, B, N, with_b, d_y);
    VK_TRY(launch_check(ctx));
    k_ipa_transcript_begin<<<ceil_div_u64(B, 64), 64, 0, s>>>(st, B, d_C, d_points, pre, mode, d_prefix_each, each_len);
    VK_TRY(launch_check(ctx));
    for (uint32_t r = 0; r < rounds; ++r) {
        uint32_t m = N >> (r + 1);
        VK_TRY(launch_fold_scalars(ctx, st, B, N, r ? 2 * m : 0, m, T, with_b, nullptr));
        VK_TRY(fixed_base_msm(ctx, k, sc, T, 2 * B, m, with_b ? k.n : 0xffffffffu, lr));
        if (B <= 64)
            k_ipa_challenge<true><<<(uint32_t)B, 32, 0, s>>>(st, B, r, rounds, d_L, d_R);
        else
            k_ipa_challenge<false><<<ceil_div_u64(B, 64), 64, 0, s>>>(st, B, r, rounds, d_L, d_R);
        VK_TRY(launch_check(ctx));
    }
    return launch_fold_scalars(ctx, st, B, N, 1, 0, T, with_b, d_tip);
}


// ---------------------------------------------------------------------------------------------------
// I3: low_level_verify_ipa (ipa/mod.rs:321-360).  The reference folds the commitment
//     c <- L_r + x_r c + x_r^2 R_r   (:348)   and compares with  tip * <G, s> + (w tip <b, s>) Q   (:355-358).
// Unrolled, that is ONE identity over fixed bases (window tables) and 1 + 2 log2 N variable points:
//     sum_i (tip s_i) G_i + (w tip <b,s> - w y prod x) Q  ==  (prod x) C + sum_r (prod_{j>r} x_j) L_r
//                                                               + sum_r (x_r^2 prod_{j>r} x_j) R_r
// with s_i = prod_r (x_r if bit (rounds-1-r) of i is 0), exactly the doubling loop at :349-353.
// ---------------------------------------------------------------------------------------------------
// Warp per proof: lane 0 replays the transcript (the challenges are a sequential hash chain); then every lane builds a
// contiguous block of s — the first log2(32) rounds select the block (one product per zero bit of the lane index), the
// remaining rounds are the reference's doubling loop (:349-353) inside the block — and its share of <b, s>.
__global__ void __launch_bounds__(128) k_ipa_verify_scalars(uint64_t B, uint32_t N, uint32_t rounds, const affine_t* __restrict__ C,
                                                            const fp_t* __restrict__ points, const fp_t* __restrict__ y,
                                                            const fp_t* __restrict__ tip, const affine_t* __restrict__ L,
                                                            const affine_t* __restrict__ R, const fp_t* __restrict__ b, TrPrefix pre,
                                                            int mode /* 0: verify_point, 1: verify_commitment_proof (:238-265: no point, no y, no Q) */,
                                                            fp_t* __restrict__ fixed_sc /*[B][N+1]*/, fp_t* __restrict__ var_sc /*[B][1+2 rounds]*/) {
    __shared__ fp_t xs_sh[4][17];  // per warp: x_0 .. x_{rounds-1}, then w
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t p = (uint64_t)blockIdx.x * 4 + warp;
    if (p >= B) return;  // (whole warps leave; only __syncwarp below)
    fp_t* fs = fixed_sc + p * (N + (mode == 0 ? 1 : 0));  // (mode 1: no Q term, N scalars per proof)
    fp_t* vs = var_sc + p * (1 + 2 * rounds);
    const fp_t yy = mode == 0 ? fp_load(y + p) : fp_zero<S>(), tp = fp_load(tip + p);
    if (lane == 0) {
        transcript_t t;
        tr_begin(t, pre);
        affine_t c;
        c.x = fp_load(&C[p].x);
        c.y = fp_load(&C[p].y);
        tr_append_point(t, c, "C");
        if (mode == 0) {
            tr_append_fr(t, fp_load(points + p), "input point");
            tr_append_fr(t, yy, "output point");
            xs_sh[warp][16] = tr_digest(t, "w");
        } else {
            tr_digest(t, "x");                      // ipa/mod.rs:249: drawn and overwritten by the first round's challenge
            xs_sh[warp][16] = fp_zero<S>();         // w = 0: the Q term vanishes
        }
        for (uint32_t r = 0; r < rounds; ++r) {
            affine_t l, rr;
            l.x = fp_load(&L[p * rounds + r].x);
            l.y = fp_load(&L[p * rounds + r].y);
            rr.x = fp_load(&R[p * rounds + r].x);
            rr.y = fp_load(&R[p * rounds + r].y);
            tr_append_point(t, l, "L");
            tr_append_point(t, rr, "R");
            xs_sh[warp][r] = tr_digest(t, "x");
        }
    }
    __syncwarp();
    // s: lanes_used blocks of `per` consecutive entries
    uint32_t top = rounds < 5 ? rounds : 5;
    const uint32_t lanes_used = 1u << top, per = N >> top;
    fp_t cb = fp_zero<S>();
    if (lane < lanes_used) {
        fp_t v0 = fp_one<S>();
        for (uint32_t r = 0; r < top; ++r)
            if (((lane >> (top - 1 - r)) & 1) == 0) v0 = fp_mul_ni<S>(v0, xs_sh[warp][r]);
        fp_t* blk = fs + (uint64_t)lane * per;
        fp_store(blk, v0);
        for (uint32_t r = top; r < rounds; ++r) {
            const fp_t x = xs_sh[warp][r];
            uint32_t len = 1u << (r - top);
            for (uint32_t i = len; i-- > 0;) {
                fp_t v = fp_load(blk + i);
                fp_store(blk + 2 * i + 1, v);
                fp_store(blk + 2 * i, fp_mul_ni<S>(v, x));
            }
        }
        // <b, s>, then s_i <- tip * s_i
        const fp_t* bb = b + p * N + (uint64_t)lane * per;
        for (uint32_t i = 0; i < per; ++i) {
            fp_t si = fp_load(blk + i);
            if (mode == 0) cb = fp_add<S>(cb, fp_mul_ni<S>(fp_load(bb + i), si));
            fp_store(blk + i, fp_mul_ni<S>(si, tp));
        }
    }
    cb = warp_sum_fr(cb);
    if (lane != 0) return;
    // suffix products prod_{j>r} x_j
    fp_t suf = fp_one<S>();
    for (uint32_t r = rounds; r-- > 0;) {
        const fp_t x = xs_sh[warp][r];
        fp_store(vs + 1 + r, suf);                                            // L_r
        fp_store(vs + 1 + rounds + r, fp_mul_ni<S>(fp_mul_ni<S>(x, x), suf));  // R_r
        suf = fp_mul_ni<S>(suf, x);
    }
    fp_store(vs, suf);  // C
    const fp_t w = xs_sh[warp][16];
    if (mode != 0) return;
    fp_t qs = fp_sub<S>(fp_mul_ni<S>(fp_mul_ni<S>(w, tp), cb), fp_mul_ni<S>(fp_mul_ni<S>(w, yy), suf));
    fp_store(fs + N, qs);
}

// variable-base scalar multiplications over the proof's own points C, L_r, R_r, 4-bit windows (warp_util.cuh).
// QUAD = false: a thread per (proof, point), for batches that fill the GPU; QUAD = true: four lanes per (proof, point) —
// a handful of verifications last as long as ONE scalar multiplication, whose dependent chain var_mul_quad shortens.
template <bool QUAD>
__global__ void __launch_bounds__(128) k_var_scalar_mul(uint64_t B, uint32_t rounds, const affine_t* __restrict__ C,
                                                        const affine_t* __restrict__ L, const affine_t* __restrict__ R,
                                                        const fp_t* __restrict__ var_sc, xyzz_t* __restrict__ out) {
    const uint32_t per = 1 + 2 * rounds;
    uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> (QUAD ? 2 : 0);
    const bool live = t < B * per;
    if (!QUAD && !live) return;  // (QUAD: no early return, the quad routines shuffle across the whole warp)
    affine_t P = affine_inf();
    fp_t k = fp_zero<S>();
    if (live) {
        uint64_t p = t / per;
        uint32_t j = (uint32_t)(t % per);
        const affine_t* src = j == 0 ? C + p : (j <= rounds ? L + p * rounds + (j - 1) : R + p * rounds + (j - 1 - rounds));
        P.x = fp_load(&src->x);
        P.y = fp_load(&src->y);
        k = fp_from_mont<S>(fp_load(var_sc + t));
    }
    if (QUAD) {
        xyzz_t acc = var_mul_quad(P, k);
        if (live && (threadIdx.x & 3) == 0) out[t] = acc;
    } else {
        out[t] = var_mul_windowed(P, k);
    }
}

// warp per proof: the 1 + 2 log2 N variable-base products are summed by a shuffle tree (per <= 33: one pass of 32 + a tail)
__global__ void __launch_bounds__(128) k_ipa_verify_final(uint64_t B, uint32_t per, const xyzz_t* __restrict__ fixed,
                                                          const xyzz_t* __restrict__ var, int32_t* __restrict__ ok) {
    uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (p >= B) return;  // (whole warps)
    xyzz_t v = xyzz_inf();
    for (uint32_t j = lane; j < per; j += 32) v = xyzz_add_ni(v, var[p * per + j]);
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) v = xyzz_add_pair(v, off);
    if (lane != 0) return;
    xyzz_t f = fixed[p];
    bool fi = xyzz_is_inf(f), vi = xyzz_is_inf(v);
    bool eq;
    if (fi || vi) {
        eq = fi && vi;
    } else {
        eq = fp_eq(fp_mul_ni<Q>(f.x, v.zz), fp_mul_ni<Q>(v.x, f.zz)) && fp_eq(fp_mul_ni<Q>(f.y, v.zzz), fp_mul_ni<Q>(v.y, f.zzz));
    }
    ok[p] = eq ? 1 : 0;
}

int32_t ipa_verify_core(vkzg_ctx* ctx, const Key& k, int mode, const fp_t* d_points, const affine_t* d_C, uint64_t B, const uint8_t* prefix,
                        uint32_t prefix_len, const char* dst, const affine_t* d_L, const affine_t* d_R, const fp_t* d_tip,
                        const fp_t* d_y, int32_t* d_ok) {
    if (B == 0) return VKZG_OK;
    const uint32_t N = k.n;
    if (N < 2 || (N & (N - 1)) || (mode == 0 && !k.has_q)) return VKZG_ERR_UNSUPPORTED;
    TrPrefix pre;
    VK_TRY(make_prefix(pre, prefix, prefix_len, dst));
    const uint32_t rounds = k.log2n, per = 1 + 2 * rounds;
    if (rounds > 16) return VKZG_ERR_UNSUPPORTED;
    DevBuf<fp_t> b, fs, vs;
    DevBuf<xyzz_t> F, V;
    VK_TRY(b.alloc(ctx, mode == 0 ? B * N : 1));
    VK_TRY(fs.alloc(ctx, B * (N + 1)));
    VK_TRY(vs.alloc(ctx, B * per));
    VK_TRY(F.alloc(ctx, B));
    VK_TRY(V.alloc(ctx, B * per));
    cudaStream_t s = ctx->stream;
    if (mode == 0) VK_TRY(barycentric_batch(ctx, k, d_points, B, b));
    k_ipa_verify_scalars<<<ceil_div_u64(B, 4), 128, 0, s>>>(B, N, rounds, d_C, d_points, d_y, d_tip, d_L, d_R, b, pre, mode, fs, vs);
    VK_TRY(launch_check(ctx));
    // (mode 1: the (N+1)-th scalar is zero and a key without Q has no row for it: N terms)
    VK_TRY(fixed_base_msm(ctx, k, fs, mode == 0 ? N + 1 : N, B, 0, 0xffffffffu, F));
    if (B * per <= (uint64_t)ctx->sm_count * 64)
        k_var_scalar_mul<true><<<ceil_div_u64(B * per * 4, 128), 128, 0, s>>>(B, rounds, d_C, d_L, d_R, vs, V);
    else
        k_var_scalar_mul<false><<<ceil_div_u64(B * per, 128), 128, 0, s>>>(B, rounds, d_C, d_L, d_R, vs, V);
    VK_TRY(launch_check(ctx));
    k_ipa_verify_final<<<ceil_div_u64(B * 32, 128), 128, 0, s>>>(B, per, F, V, d_ok);
    return launch_check(ctx);
}

}  // namespace vk

using namespace vk;

extern "C" {

int32_t vkzg_barycentric_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* points, uint64_t B, vkzg_fr* out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!points || !out))) return VKZG_ERR_ARG;
    DevBuf<fp_t> dp, dout;
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(dout.alloc(ctx, B * k->n));
    VK_TRY(barycentric_batch(ctx, *k, dp, B, dout));
    VK_TRY(download(ctx, out, dout.p, B * k->n));
    return stream_sync(ctx);
}

int32_t vkzg_ipa_prove_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_a, const vkzg_fr* d_points,
                                 const vkzg_g1_affine* d_commitments, uint64_t B, const uint8_t* prefix, uint32_t prefix_len,
                                 const char* dst, vkzg_g1_affine* d_L, vkzg_g1_affine* d_R, vkzg_fr* d_tip, vkzg_fr* d_y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (B && (!d_a || !d_points || !d_commitments || !d_L || !d_R || !d_tip || !d_y)) return VKZG_ERR_ARG;
    return ipa_prove_core(ctx, *k, 0, k->n, (const fp_t*)d_a, (const fp_t*)d_points, (const affine_t*)d_commitments, B, prefix,
                          prefix_len, dst, (affine_t*)d_L, (affine_t*)d_R, (fp_t*)d_tip, (fp_t*)d_y);
}

int32_t vkzg_ipa_prove_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* a, const vkzg_fr* points,
                             const vkzg_g1_affine* commitments, uint64_t B, const uint8_t* prefix, uint32_t prefix_len,
                             const char* dst, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (B && (!a || !points || !commitments || !L || !R || !tip || !y)) return VKZG_ERR_ARG;
    uint32_t N = k->n, rounds = k->log2n;
    if (B == 0) return VKZG_OK;
    DevBuf<fp_t> da, dp, dtip, dy;
    DevBuf<affine_t> dc, dL, dR;
    VK_TRY(da.alloc(ctx, B * N));
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(upload(ctx, dc, commitments, B));
    VK_TRY(dL.alloc(ctx, B * rounds));
    VK_TRY(dR.alloc(ctx, B * rounds));
    VK_TRY(dtip.alloc(ctx, B));
    VK_TRY(dy.alloc(ctx, B));
    // One chunk: the per-round challenge / fold kernels are latency-bound (their duration does not shrink with the
    // batch), so proving in pieces to overlap the upload costs more than the ~5 % of PCIe time it could hide (measured).
    ChunkedUpload up(ctx);
    VK_TRY(up.init());
    const uint64_t chunk = B;
    for (uint64_t b0 = 0; b0 < B; b0 += chunk) {
        uint64_t nb = B - b0 < chunk ? B - b0 : chunk;
        VK_TRY(up.copy(da.p + b0 * N, (const fp_t*)a + b0 * N, nb * N * sizeof(fp_t)));
        VK_TRY(up.publish());
        VK_TRY(ipa_prove_core(ctx, *k, 0, N, da.p + b0 * N, dp.p + b0, dc.p + b0, nb, prefix, prefix_len, dst, dL.p + b0 * rounds,
                              dR.p + b0 * rounds, dtip.p + b0, dy.p + b0));
    }
    VK_TRY(download(ctx, L, dL.p, B * rounds));
    VK_TRY(download(ctx, R, dR.p, B * rounds));
    VK_TRY(download(ctx, tip, dtip.p, B));
    VK_TRY(download(ctx, y, dy.p, B));
    return stream_sync(ctx);
}

// commit + open in one call: the rows cross PCIe once (chunked, overlapped with the commit kernels), the commitments
// never leave the device between the two steps
int32_t vkzg_ipa_commit_prove_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* a, const vkzg_fr* points, uint64_t B,
                                    vkzg_g1_affine* commitments, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* y) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || !k->has_q) return VKZG_ERR_ARG;
    if (B && (!a || !points || !commitments || !L || !R || !tip || !y)) return VKZG_ERR_ARG;
    uint32_t N = k->n, rounds = k->log2n;
    if (B == 0) return VKZG_OK;
    DevBuf<fp_t> da, dp, dtip, dy;
    DevBuf<affine_t> dc, dL, dR;
    DevBuf<xyzz_t> acc;
    VK_TRY(da.alloc(ctx, B * N));
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(dc.alloc(ctx, B));
    VK_TRY(acc.alloc(ctx, B));
    VK_TRY(dL.alloc(ctx, B * rounds));
    VK_TRY(dR.alloc(ctx, B * rounds));
    VK_TRY(dtip.alloc(ctx, B));
    VK_TRY(dy.alloc(ctx, B));
    ChunkedUpload up(ctx);
    VK_TRY(up.init());
    for (uint64_t b0 = 0, nb = 0; b0 < B; b0 += nb) {
        nb = pipeline_piece(B, b0);
        VK_TRY(up.copy(da.p + b0 * N, (const fp_t*)a + b0 * N, nb * N * sizeof(fp_t)));
        VK_TRY(up.publish());
        VK_TRY(fixed_base_msm(ctx, *k, da.p + b0 * N, N, nb, 0, 0xffffffffu, acc.p + b0));
    }
    VK_TRY(normalize_points(ctx, acc, B, dc));
    VK_TRY(ipa_prove_core(ctx, *k, 0, N, da, dp, dc, B, nullptr, 0, "ipa", dL, dR, dtip, dy));
    VK_TRY(download(ctx, commitments, dc.p, B));
    VK_TRY(download(ctx, L, dL.p, B * rounds));
    VK_TRY(download(ctx, R, dR.p, B * rounds));
    VK_TRY(download(ctx, tip, dtip.p, B));
    VK_TRY(download(ctx, y, dy.p, B));
    return stream_sync(ctx);
}

int32_t vkzg_ipa_prove_commitment_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* a, const vkzg_g1_affine* commitments,
                                        uint64_t B, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (B && (!a || !commitments || !L || !R || !tip)) return VKZG_ERR_ARG;
    uint32_t N = k->n, rounds = k->log2n;
    if (N & (N - 1)) return VKZG_ERR_UNSUPPORTED;
    DevBuf<fp_t> da, dtip, dy;
    DevBuf<affine_t> dc, dL, dR;
    VK_TRY(upload(ctx, da, a, B * N));
    VK_TRY(upload(ctx, dc, commitments, B));
    VK_TRY(dL.alloc(ctx, B * rounds));
    VK_TRY(dR.alloc(ctx, B * rounds));
    VK_TRY(dtip.alloc(ctx, B));
    VK_TRY(dy.alloc(ctx, B));
    VK_TRY(ipa_prove_core(ctx, *k, 1, N, da, nullptr, dc, B, nullptr, 0, "ipa", dL, dR, dtip, dy));
    VK_TRY(download(ctx, L, dL.p, B * rounds));
    VK_TRY(download(ctx, R, dR.p, B * rounds));
    VK_TRY(download(ctx, tip, dtip.p, B));
    return stream_sync(ctx);
}

int32_t vkzg_ipa_verify_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* points, const vkzg_g1_affine* commitments, uint64_t B,
                              const uint8_t* prefix, uint32_t prefix_len, const char* dst, const vkzg_g1_affine* L,
                              const vkzg_g1_affine* R, const vkzg_fr* tip, const vkzg_fr* y, int32_t* ok) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (B && (!points || !commitments || !L || !R || !tip || !y || !ok)) return VKZG_ERR_ARG;
    uint32_t rounds = k->log2n;
    DevBuf<fp_t> dp, dtip, dy;
    DevBuf<affine_t> dc, dL, dR;
    DevBuf<int32_t> dok;
    VK_TRY(upload(ctx, dp, points, B));
    VK_TRY(upload(ctx, dc, commitments, B));
    VK_TRY(upload(ctx, dL, L, B * rounds));
    VK_TRY(upload(ctx, dR, R, B * rounds));
    VK_TRY(upload(ctx, dtip, tip, B));
    VK_TRY(upload(ctx, dy, y, B));
    VK_TRY(dok.alloc(ctx, B));
    VK_TRY(ipa_verify_core(ctx, *k, 0, dp, dc, B, prefix, prefix_len, dst, dL, dR, dtip, dy, dok));
    VK_TRY(download(ctx, ok, dok.p, B));
    return stream_sync(ctx);
}

int32_t vkzg_ipa_verify_commitment_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_g1_affine* commitments, uint64_t B,
                                         const vkzg_g1_affine* L, const vkzg_g1_affine* R, const vkzg_fr* tip, int32_t* ok) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (B && (!commitments || !L || !R || !tip || !ok)) return VKZG_ERR_ARG;
    uint32_t rounds = k->log2n;
    DevBuf<fp_t> dtip;
    DevBuf<affine_t> dc, dL, dR;
    DevBuf<int32_t> dok;
    VK_TRY(upload(ctx, dc, commitments, B));
    VK_TRY(upload(ctx, dL, L, B * rounds));
    VK_TRY(upload(ctx, dR, R, B * rounds));
    VK_TRY(upload(ctx, dtip, tip, B));
    VK_TRY(dok.alloc(ctx, B));
    VK_TRY(ipa_verify_core(ctx, *k, 1, nullptr, dc, B, nullptr, 0, "ipa", dL, dR, dtip, nullptr, dok));
    VK_TRY(download(ctx, ok, dok.p, B));
    return stream_sync(ctx);
}

}  // extern "C"
