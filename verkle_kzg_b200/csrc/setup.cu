// KZG::setup (kzg/mod.rs:115-124), the "next" row after the hot path: the Lagrange-form SRS is the
// GROUP inverse FFT of the powers-of-tau points,  L_j = (1/n) sum_i [tau^i]G * w^(-ij),  over the radix-2 domain of size
// n = next_pow2(max_items) (inputs beyond max_items are the identity, as ark-poly's ifft zero-pads).
//   k_gfft_stage   four lanes per butterfly of a decimation-in-frequency stage: (u, v) -> (u + v, (u - v) * w^-k); the
//                  twiddle multiplication is a 4-bit-window scalar multiplication whose products are spread over the quad
//                  (a stage is n/2 chains of ~320 dependent point operations: latency, not throughput, bounds it)
//   k_gfft_finish  bit-reversal + multiplication by 1/n, then the common batched normalisation
// kzg_point_generator.rs:32-43 (G * tau^i) is a width-1 fixed-base batch against a one-base key: vkzg_kzg_powers.
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

// k * P for an XYZZ point held by all four lanes of a quad (warp_util.cuh: the products of one point operation are spread
// over the quad): fixed 4-bit windows over P .. 15P — 252 doublings + <= 64 additions, each ~40 % shorter than a lone
// thread's.  All 32 lanes of the warp must call it.
static __device__ __noinline__ xyzz_t xyzz_scalar_mul_quad(const xyzz_t P, const fp_t k_canon) {
    xyzz_t T[15];
    T[0] = P;
#pragma unroll 1
    for (int d = 2; d <= 15; ++d) T[d - 1] = (d & 1) ? xyzz_add_quad(T[d - 2], T[0]) : xyzz_dbl_quad(T[d / 2 - 1]);
    xyzz_t acc = xyzz_inf();
#pragma unroll 1
    for (int w = 63; w >= 0; --w) {
        if (w != 63) {
#pragma unroll 1
            for (int j = 0; j < 4; ++j) acc = xyzz_dbl_quad(acc);
        }
        uint32_t limb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i == (w >> 3)) limb = k_canon.l[i];
        const uint32_t d = (limb >> (4 * (w & 7))) & 15;
        if (__any_sync(0xffffffffu, d != 0)) {
            xyzz_t t = xyzz_add_quad(acc, T[d ? d - 1 : 0]);
            if (d) acc = t;
        }
    }
    return acc;
}

__global__ void __launch_bounds__(128) k_gfft_load(const affine_t* __restrict__ in, uint32_t m, uint32_t n, xyzz_t* __restrict__ x) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_t p = xyzz_inf();
    if (i < m) {
        affine_t a;
        a.x = fp_load(&in[i].x);
        a.y = fp_load(&in[i].y);
        p = xyzz_from_affine(a);
    }
    x[i] = p;
}

// half = h: butterflies (s + k, s + k + h) for every block start s (multiple of 2h) and k < h; FOUR lanes per butterfly
__global__ void __launch_bounds__(128) k_gfft_stage(xyzz_t* __restrict__ x, uint32_t n, uint32_t h, const fp_t* __restrict__ omega_inv) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t t = gt >> 2, q = gt & 3;
    const bool live = t < n / 2;
    const uint32_t tt = live ? t : 0;
    const uint32_t k = tt % h, s = (tt / h) * 2 * h;
    const xyzz_t u = x[s + k], v = x[s + k + h];
    const xyzz_t sum = xyzz_add_quad(u, v);
    xyzz_t dif = xyzz_add_quad(u, xyzz_neg(v));
    fp_t w = fp_zero<S>();
    w.l[0] = 1;  // canonical 1: k = 0 multiplies by one (the quad form needs every lane of the warp in the loop)
    if (k) w = fp_from_mont<S>(fp_load_ro(omega_inv + (size_t)k * (n / (2 * h))));
    if (__any_sync(0xffffffffu, k != 0)) {
        const xyzz_t m = xyzz_scalar_mul_quad(dif, w);
        if (k) dif = m;
    }
    if (live && q == 0) {
        x[s + k] = sum;
        x[s + k + h] = dif;
    }
}

// bit-reversal + multiplication by 1/n, four lanes per point
__global__ void __launch_bounds__(128) k_gfft_finish(const xyzz_t* __restrict__ x, uint32_t n, uint32_t lg, fp_t n_inv, xyzz_t* __restrict__ out) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = gt >> 2, q = gt & 3;
    const bool live = i < n;
    const uint32_t ii = live ? i : 0;
    const uint32_t r = lg ? (__brev(ii) >> (32 - lg)) : 0;
    const xyzz_t m = xyzz_scalar_mul_quad(x[ii], fp_from_mont<S>(n_inv));
    if (live && q == 0) out[r] = m;
}

int32_t kzg_setup_core(vkzg_ctx* ctx, const affine_t* d_powers, uint32_t m, affine_t* d_out) {
    uint32_t lg = 0;
    while ((1u << lg) < m) ++lg;
    const uint32_t n = 1u << lg;
    const DomainTables* dt;
    VK_TRY(domain_for(ctx, lg, dt));
    DevBuf<xyzz_t> x, y;
    VK_TRY(x.alloc(ctx, n));
    VK_TRY(y.alloc(ctx, n));
    cudaStream_t s = ctx->stream;
    k_gfft_load<<<ceil_div_u64(n, 128), 128, 0, s>>>(d_powers, m, n, x);
    VK_TRY(launch_check(ctx));
    for (uint32_t h = n / 2; h >= 1; h /= 2) {
        k_gfft_stage<<<ceil_div_u64((uint64_t)n * 2, 128), 128, 0, s>>>(x, n, h, dt->omega_inv);
        VK_TRY(launch_check(ctx));
    }
    k_gfft_finish<<<ceil_div_u64((uint64_t)n * 4, 128), 128, 0, s>>>(x, n, lg, dt->n_inv, y);
    VK_TRY(launch_check(ctx));
    return normalize_points(ctx, y, n, d_out);
}

__global__ void __launch_bounds__(128) k_tau_powers(fp_t tau, uint64_t m, fp_t* __restrict__ out) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    fp_t acc = fp_one<S>();
    for (int b = 63 - __clzll(q | 1); b >= 0; --b) {
        acc = fp_mul_ni<S>(acc, acc);
        if ((q >> b) & 1) acc = fp_mul_ni<S>(acc, tau);
    }
    fp_store(out + q, acc);
}

// KZG::setup when the generator's secret is at hand (kzg/mod.rs:115-124 reads `gen.secret()` for the G2 element anyway):
// the inverse FFT of (tau^i G)_{i < m}, zero-padded to n, is  L_j = s_j G  with the truncated geometric sum
//     s_j = (1/n) sum_{i<m} (tau w^-j)^i = (1/n) ((tau w^-j)^m - 1) / (tau w^-j - 1)      (m/n when tau w^-j = 1),
// so the Lagrange SRS is n scalars (one shared inversion per warp) and ONE batch of width-1 fixed-base jobs — the same
// canonical affine points as the group FFT above, at a fraction of its dependent point operations.
__global__ void __launch_bounds__(128) k_lagrange_scalars(fp_t tau, uint32_t m, uint32_t n, const fp_t* __restrict__ omega_inv, fp_t n_inv,
                                                          fp_t* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim is a multiple of 32: whole warps reach the shuffles
    const bool live = j < n;
    fp_t x = fp_one<S>();
    if (live) x = fp_mul_ni<S>(tau, fp_load_ro(omega_inv + j));
    fp_t den = fp_sub<S>(x, fp_one<S>());
    const bool unit = fp_is_zero(den);
    fp_t pw = fp_one<S>();
#pragma unroll 1
    for (int b = 31 - __clz(m | 1); b >= 0; --b) {
        pw = fp_mul_ni<S>(pw, pw);
        if ((m >> b) & 1) pw = fp_mul_ni<S>(pw, x);
    }
    const fp_t inv = warp_inverse_of_lane_products(unit || !live ? fp_one<S>() : den);
    if (!live) return;
    fp_t sc;
    if (unit) {
        fp_t mm = fp_zero<S>();
        mm.l[0] = m;
        sc = fp_to_mont<S>(mm);
    } else {
        sc = fp_mul_ni<S>(fp_sub<S>(pw, fp_one<S>()), inv);
    }
    fp_store(out + j, fp_mul_ni<S>(sc, n_inv));
}

}  // namespace vk

using namespace vk;

extern "C" {

int32_t vkzg_kzg_setup_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_powers, uint32_t m, vkzg_g1_affine* d_lagrange) {
    VK_TRY(ctx_check(ctx));
    if (!m || !d_powers || !d_lagrange) return VKZG_ERR_ARG;
    if (m > (1u << 24)) return VKZG_ERR_RANGE;
    return kzg_setup_core(ctx, (const affine_t*)d_powers, m, (affine_t*)d_lagrange);
}

int32_t vkzg_kzg_setup(vkzg_ctx* ctx, const vkzg_g1_affine* powers, uint32_t m, vkzg_g1_affine* lagrange) {
    VK_TRY(ctx_check(ctx));
    if (!m || !powers || !lagrange) return VKZG_ERR_ARG;
    if (m > (1u << 24)) return VKZG_ERR_RANGE;
    uint32_t n = 1;
    while (n < m) n <<= 1;
    DevBuf<affine_t> dp, dl;
    VK_TRY(upload(ctx, dp, powers, m));
    VK_TRY(dl.alloc(ctx, n));
    VK_TRY(kzg_setup_core(ctx, dp, m, dl));
    VK_TRY(download(ctx, lagrange, dl.p, n));
    return stream_sync(ctx);
}

int32_t vkzg_kzg_setup_from_secret(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* tau, uint32_t m, vkzg_g1_affine* lagrange) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || !tau || !m || !lagrange) return VKZG_ERR_ARG;
    if (m > (1u << 24)) return VKZG_ERR_RANGE;
    uint32_t lg = 0;
    while ((1u << lg) < m) ++lg;
    const uint32_t n = 1u << lg;
    const DomainTables* dt;
    VK_TRY(domain_for(ctx, lg, dt));
    DevBuf<fp_t> sc;
    DevBuf<xyzz_t> acc;
    DevBuf<affine_t> dout;
    VK_TRY(sc.alloc(ctx, n));
    VK_TRY(acc.alloc(ctx, n));
    VK_TRY(dout.alloc(ctx, n));
    fp_t t;
    memcpy(&t, tau, sizeof(t));
    k_lagrange_scalars<<<ceil_div_u64(n, 128), 128, 0, ctx->stream>>>(t, m, n, dt->omega_inv, dt->n_inv, sc);
    VK_TRY(launch_check(ctx));
    VK_TRY(fixed_base_msm(ctx, *k, sc, 1, n, 0, 0xffffffffu, acc));
    VK_TRY(normalize_points(ctx, acc, n, dout));
    VK_TRY(download(ctx, lagrange, dout.p, n));
    return stream_sync(ctx);
}

// KZGRandomPointGenerator::gen (kzg_point_generator.rs:32-43): out[i] = tau^i * G for i < m, with `key_id` a
// VKZG_KEY_WINDOW key whose first base is the generator G
int32_t vkzg_kzg_powers(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* tau, uint32_t m, vkzg_g1_affine* out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || !tau || !m || !out) return VKZG_ERR_ARG;
    DevBuf<fp_t> sc;
    DevBuf<xyzz_t> acc;
    DevBuf<affine_t> dout;
    VK_TRY(sc.alloc(ctx, m));
    VK_TRY(acc.alloc(ctx, m));
    VK_TRY(dout.alloc(ctx, m));
    fp_t t;
    memcpy(&t, tau, sizeof(t));
    k_tau_powers<<<ceil_div_u64(m, 128), 128, 0, ctx->stream>>>(t, m, sc);
    VK_TRY(launch_check(ctx));
    VK_TRY(fixed_base_msm(ctx, *k, sc, 1, m, 0, 0xffffffffu, acc));
    VK_TRY(normalize_points(ctx, acc, m, dout));
    VK_TRY(download(ctx, out, dout.p, m));
    return stream_sync(ctx);
}

}  // extern "C"
