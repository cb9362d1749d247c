// KZG::setup (kzg/mod.rs:115-124), the "next" row after the hot path: the Lagrange-form SRS is the
// GROUP inverse FFT of the powers-of-tau points,  L_j = (1/n) sum_i [tau^i]G * w^(-ij),  over the radix-2 domain of size
// n = next_pow2(max_items) (inputs beyond max_items are the identity, as ark-poly's ifft zero-pads).
//   k_gfft_stage   one thread per butterfly of a decimation-in-frequency stage: (u, v) -> (u + v, (u - v) * w^-k), the
//                  twiddle multiplication a 254-bit double-and-add on XYZZ points
//   k_gfft_finish  bit-reversal + multiplication by 1/n, then the common batched normalisation
// kzg_point_generator.rs:32-43 (G * tau^i) is a width-1 fixed-base batch against a one-base key: vkzg_kzg_powers.
#include "vk_common.cuh"

namespace vk {

__device__ __forceinline__ xyzz_t xyzz_scalar_mul(const xyzz_t& p, const fp_t& k_canon) {
    xyzz_t acc = xyzz_inf();
    int top = -1;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (k_canon.l[i]) top = 32 * i + 31 - __clz(k_canon.l[i]);
#pragma unroll 1
    for (int bit = top; bit >= 0; --bit) {
        acc = xyzz_dbl_ni(acc);
        uint32_t limb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i == (bit >> 5)) limb = k_canon.l[i];
        if ((limb >> (bit & 31)) & 1) acc = xyzz_add_ni(acc, p);
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

// half = h: butterflies (s + k, s + k + h) for every block start s (multiple of 2h) and k < h
__global__ void __launch_bounds__(128) k_gfft_stage(xyzz_t* __restrict__ x, uint32_t n, uint32_t h, const fp_t* __restrict__ omega_inv) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    uint32_t k = t % h, s = (t / h) * 2 * h;
    xyzz_t u = x[s + k], v = x[s + k + h];
    xyzz_t sum = xyzz_add_ni(u, v);
    xyzz_t dif = xyzz_add_ni(u, xyzz_neg(v));
    if (k) {
        fp_t w = fp_from_mont<S>(fp_load_ro(omega_inv + (size_t)k * (n / (2 * h))));
        dif = xyzz_scalar_mul(dif, w);
    }
    x[s + k] = sum;
    x[s + k + h] = dif;
}

__global__ void __launch_bounds__(128) k_gfft_finish(const xyzz_t* __restrict__ x, uint32_t n, uint32_t lg, fp_t n_inv, xyzz_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = lg ? (__brev(i) >> (32 - lg)) : 0;
    out[r] = xyzz_scalar_mul(x[i], fp_from_mont<S>(n_inv));
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
        k_gfft_stage<<<ceil_div_u64(n / 2, 128), 128, 0, s>>>(x, n, h, dt->omega_inv);
        VK_TRY(launch_check(ctx));
    }
    k_gfft_finish<<<ceil_div_u64(n, 128), 128, 0, s>>>(x, n, lg, dt->n_inv, y);
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
