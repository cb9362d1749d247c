// Key-load kernels: fixed-base signed-window tables (width-N commits / IPA / KZG opens), the 2^(c*w)
// multiples used by the shared-bucket Pippenger MSM, domain constants, and the batched XYZZ -> affine
// normalisation every point result goes through.
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

// -------------------------------------------------------------------------------------------------
// XYZZ -> affine by Montgomery's trick: every lane multiplies the zzz of its K points (prefix products parked in the x
// coordinate of the OUTPUT slots), the 32 lane products are inverted together with ONE inversion per warp
// (warp_inverse_of_lane_products_t: all lanes invert the same grand total, so the data-dependent binary inversion runs
// without divergence), i.e. 32 K points share one inversion.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_normalize(const xyzz_t* __restrict__ in, uint64_t n, uint32_t K, affine_t* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t first = t * K;
    // (no early return: the warp-wide inversion below needs all 32 lanes; lanes past the end carry the product 1)
    uint32_t cnt = first >= n ? 0 : (uint32_t)(n - first < K ? n - first : K);
    fp_t run = fp_one<Q>();
#pragma unroll 1
    for (uint32_t j = 0; j < cnt; ++j) {
        fp_store(&out[first + j].x, run);
        fp_t z = fp_load(&in[first + j].zzz);
        if (!fp_is_zero(z)) run = fp_mul_ni<Q>(run, z);
    }
    // one inversion per warp, on a value all lanes share (uniform control flow), instead of 32 divergent ones
    fp_t inv = warp_inverse_of_lane_products_t<Q>(run);
#pragma unroll 1
    for (uint32_t j = cnt; j-- > 0;) {
        xyzz_t p;
        p.x = fp_load(&in[first + j].x);
        p.y = fp_load(&in[first + j].y);
        p.zz = fp_load(&in[first + j].zz);
        p.zzz = fp_load(&in[first + j].zzz);
        affine_t a = affine_inf();
        if (!fp_is_zero(p.zzz)) {
            fp_t zinv = fp_mul_ni<Q>(inv, fp_load(&out[first + j].x));
            inv = fp_mul_ni<Q>(inv, p.zzz);
            fp_t tt = fp_mul_ni<Q>(zinv, p.zz);
            fp_t zz_inv = fp_mul_ni<Q>(tt, tt);
            a.x = fp_mul_ni<Q>(p.x, zz_inv);
            a.y = fp_mul_ni<Q>(p.y, zinv);
        }
        fp_store(&out[first + j].x, a.x);
        fp_store(&out[first + j].y, a.y);
    }
}

int32_t normalize_points(vkzg_ctx* ctx, const xyzz_t* d_in, uint64_t n, affine_t* d_out) {
    if (n == 0) return VKZG_OK;
    if ((const void*)d_in == (const void*)d_out) return VKZG_ERR_ARG;  // the output doubles as scratch
    uint32_t K = n <= (1u << 16) ? 1 : (n <= (1u << 19) ? 4 : 8);
    k_normalize<<<ceil_div_u64((n + K - 1) / K, 128), 128, 0, ctx->stream>>>(d_in, n, K, d_out);
    return launch_check(ctx);
}

// -------------------------------------------------------------------------------------------------
// Fixed-base window tables.  table[((b * W + w) << (c-1)) + (m - 1)] = m * 2^(c w) * base_b, 1 <= m <= 2^(c-1)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_wtab_seed(const affine_t* __restrict__ bases, uint32_t nb, uint32_t W, uint32_t c,
                                                  affine_t* __restrict__ table) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nb * W) return;
    uint32_t b = t / W, w = t % W;
    affine_t P = bases[b];
    xyzz_t acc = xyzz_from_affine(P);
    for (uint32_t k = 0; k < w * c; ++k) acc = xyzz_dbl_ni(acc);
    affine_t a = xyzz_to_affine(acc);
    table[(size_t)t << (c - 1)] = a;
}

// Level k: entries m in (2^k, 2^(k+1)] from entries (0, 2^k] plus B = entry 2^k; affine additions whose
// inversions are shared G at a time.
template <int G>
__global__ void __launch_bounds__(128) k_wtab_level(affine_t* __restrict__ table, uint32_t rows, uint32_t c, uint32_t k) {
    const uint32_t per = 1u << k;
    const uint32_t g = per < (uint32_t)G ? per : (uint32_t)G;
    const uint32_t groups = per / g;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < (uint64_t)rows * groups;  // (no early return: the warp-wide inversion needs all 32 lanes)
    uint32_t row = live ? (uint32_t)(t / groups) : 0, grp = live ? (uint32_t)(t % groups) : 0;
    affine_t* T = table + ((size_t)row << (c - 1));
    affine_t Bp = T[per - 1];
    bool binf = affine_is_inf(Bp);
    fp_t pre[G];
    fp_t run = fp_one<Q>();
    const uint32_t gg = live ? g : 0;
#pragma unroll 1
    for (uint32_t j = 0; j < gg; ++j) {
        uint32_t idx = grp * g + j;
        pre[j] = run;
        fp_t den;
        if (binf)
            den = fp_one<Q>();
        else if (idx + 1 == per)
            den = fp_dbl<Q>(Bp.y);
        else
            den = fp_sub<Q>(Bp.x, fp_load(&T[idx].x));
        run = fp_mul_ni<Q>(run, den);
    }
    fp_t inv = warp_inverse_of_lane_products_t<Q>(run);  // 32 x G additions share one inversion
#pragma unroll 1
    for (int j = (int)gg - 1; j >= 0; --j) {
        uint32_t idx = grp * g + j;
        affine_t r = affine_inf();
        if (!binf) {
            affine_t A;
            A.x = fp_load(&T[idx].x);
            A.y = fp_load(&T[idx].y);
            bool dbl = idx + 1 == per;
            fp_t den = dbl ? fp_dbl<Q>(Bp.y) : fp_sub<Q>(Bp.x, A.x);
            fp_t dinv = fp_mul_ni<Q>(inv, pre[j]);
            inv = fp_mul_ni<Q>(inv, den);
            fp_t num;
            if (dbl) {
                fp_t x2 = fp_mul_ni<Q>(A.x, A.x);
                num = fp_add<Q>(fp_dbl<Q>(x2), x2);
            } else {
                num = fp_sub<Q>(Bp.y, A.y);
            }
            fp_t lam = fp_mul_ni<Q>(num, dinv);
            fp_t x3 = fp_sub<Q>(fp_sub<Q>(fp_mul_ni<Q>(lam, lam), A.x), Bp.x);
            fp_t y3 = fp_sub<Q>(fp_mul_ni<Q>(lam, fp_sub<Q>(A.x, x3)), A.y);
            r.x = x3;
            r.y = y3;
        }
        fp_store(&T[idx + per].x, r.x);
        fp_store(&T[idx + per].y, r.y);
    }
}

int32_t build_window_tables(vkzg_ctx* ctx, Key& k) {
    uint32_t nb = k.n + (k.has_q ? 1 : 0);
    uint32_t rows = nb * k.W;
    k.table_points = (uint64_t)rows << (k.c - 1);
    VK_CUDA(cudaMalloc((void**)&k.table, k.table_points * sizeof(affine_t)));
    k_wtab_seed<<<ceil_div_u64(rows, 64), 64, 0, ctx->stream>>>(k.bases, nb, k.W, k.c, k.table);
    VK_TRY(launch_check(ctx));
    for (uint32_t lvl = 0; lvl + 1 < k.c; ++lvl) {
        uint32_t per = 1u << lvl;
        uint32_t g = per < 8 ? per : 8;
        uint64_t threads = (uint64_t)rows * (per / g);
        k_wtab_level<8><<<ceil_div_u64(threads, 128), 128, 0, ctx->stream>>>(k.table, rows, k.c, lvl);
        VK_TRY(launch_check(ctx));
    }
    return VKZG_OK;
}

// -------------------------------------------------------------------------------------------------
// MSM tables: table[w * n + i] = 2^(c w) * base_i
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mtab(const affine_t* __restrict__ bases, uint32_t n, uint32_t W, uint32_t c,
                                              xyzz_t* __restrict__ tmp) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t P;
    P.x = fp_load(&bases[i].x);
    P.y = fp_load(&bases[i].y);
    xyzz_t acc = xyzz_from_affine(P);
    tmp[i] = acc;
#pragma unroll 1
    for (uint32_t w = 1; w < W; ++w) {
#pragma unroll 1
        for (uint32_t b = 0; b < c; ++b) acc = xyzz_dbl_ni(acc);
        tmp[(size_t)w * n + i] = acc;
    }
}

int32_t build_msm_tables(vkzg_ctx* ctx, Key& k) {
    k.table_points = (uint64_t)k.n * k.W;
    VK_CUDA(cudaMalloc((void**)&k.table, k.table_points * sizeof(affine_t)));
    // doubling chains in XYZZ, then ONE batched normalisation of all n * W points (shared inversions)
    DevBuf<xyzz_t> tmp;
    VK_TRY(tmp.alloc(ctx, k.table_points));
    k_mtab<<<ceil_div_u64(k.n, 128), 128, 0, ctx->stream>>>(k.bases, k.n, k.W, k.c, tmp);
    VK_TRY(launch_check(ctx));
    return normalize_points(ctx, tmp, k.table_points, k.table);
}

// -------------------------------------------------------------------------------------------------
// Domain constants (ark-poly Radix2EvaluationDomain: group_gen = 5^((r-1)/size)).
// -------------------------------------------------------------------------------------------------
__device__ fp_t fr_pow_u64(fp_t base, uint64_t e) {
    fp_t acc = fp_one<S>();
    for (int b = 63; b >= 0; --b) {
        acc = fp_mul_ni<S>(acc, acc);
        if ((e >> b) & 1) acc = fp_mul_ni<S>(acc, base);
    }
    return acc;
}

__device__ fp_t fr_domain_gen(uint32_t log2n) {
    // (r - 1) >> log2n, r - 1 = 2^28 * odd
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = S::p(i);
    e[0] -= 1;
    fp_t five = fp_from_u32<S>(5);
    fp_t acc = fp_one<S>();
    for (int bit = 255; bit >= (int)log2n; --bit) {
        acc = fp_mul_ni<S>(acc, acc);
        uint32_t limb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i == (bit >> 5)) limb = e[i];
        if ((limb >> (bit & 31)) & 1) acc = fp_mul_ni<S>(acc, five);
    }
    return acc;
}

__global__ void k_domain(uint32_t log2n, fp_t* omega, fp_t* omega_inv, fp_t* diff_inv) {
    uint32_t n = 1u << log2n;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t g = fr_domain_gen(log2n);
    fp_t wi = fr_pow_u64(g, i);
    fp_t wni = fr_pow_u64(g, (n - i) & (n - 1));
    omega[i] = wi;
    omega_inv[i] = wni;
    diff_inv[i] = i == 0 ? fp_zero<S>() : fp_inv<S>(fp_sub<S>(wi, fp_one<S>()));
}

__global__ void k_small_consts(uint32_t n, fp_t* out2) {
    fp_t nm = fp_from_u32<S>(n);
    out2[0] = nm;
    out2[1] = fp_inv<S>(nm);
}

int32_t build_domain(vkzg_ctx* ctx, uint32_t log2n, uint32_t size_n, DomainTables& d) {
    uint32_t n = 1u << log2n;
    VK_CUDA(cudaMalloc((void**)&d.omega, 3 * (size_t)n * sizeof(fp_t)));
    d.omega_inv = d.omega + n;
    d.diff_inv = d.omega + 2 * (size_t)n;
    k_domain<<<ceil_div_u64(n, 64), 64, 0, ctx->stream>>>(log2n, d.omega, d.omega_inv, d.diff_inv);
    VK_TRY(launch_check(ctx));
    DevBuf<fp_t> two;
    VK_TRY(two.alloc(ctx, 2));
    k_small_consts<<<1, 1, 0, ctx->stream>>>(size_n, two);
    VK_TRY(launch_check(ctx));
    fp_t h[2];
    VK_CUDA(cudaMemcpyAsync(h, two.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    d.n_mont = h[0];
    d.n_inv = h[1];
    return VKZG_OK;
}

// cached constants of the radix-2 domain of size 2^lg
int32_t domain_for(vkzg_ctx* ctx, uint32_t lg, const DomainTables*& dt) {
    auto it = ctx->domains.find(lg);
    if (it == ctx->domains.end()) {
        DomainTables d;
        VK_TRY(build_domain(ctx, lg, 1u << lg, d));
        it = ctx->domains.emplace(lg, d).first;
    }
    dt = &it->second;
    return VKZG_OK;
}

int32_t build_domain_tables(vkzg_ctx* ctx, Key& k) {
    uint32_t lg = 0;
    while ((1u << lg) < k.n) ++lg;
    k.log2n = lg;
    k.domain_n = 1u << lg;
    return build_domain(ctx, lg, k.n, k.dom);
}


// Key-load validation: every base is the identity (0,0) or a canonical (coordinates < p) point with y^2 = x^3 + 3.
__global__ void __launch_bounds__(128) k_check_on_curve(const affine_t* __restrict__ pts, uint64_t n, uint32_t* __restrict__ bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t p;
    p.x = fp_load(&pts[i].x);
    p.y = fp_load(&pts[i].y);
    if (affine_is_inf(p)) return;
    uint32_t pl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pl[k] = Q::p(k);
    bool ok = !geq8(p.x.l, pl) && !geq8(p.y.l, pl);
    if (ok) {
        fp_t three = fp_from_u32<Q>(3);
        fp_t rhs = fp_add<Q>(fp_mul_ni<Q>(fp_mul_ni<Q>(p.x, p.x), p.x), three);
        ok = fp_eq(fp_mul_ni<Q>(p.y, p.y), rhs);
    }
    if (!ok) atomicAdd(bad, 1u);
}

int32_t check_points_on_curve(vkzg_ctx* ctx, const affine_t* d_points, uint64_t n) {
    if (n == 0) return VKZG_OK;
    DevBuf<uint32_t> bad;
    VK_TRY(bad.alloc(ctx, 1));
    VK_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(uint32_t), ctx->stream));
    k_check_on_curve<<<ceil_div_u64(n, 128), 128, 0, ctx->stream>>>(d_points, n, bad);
    VK_TRY(launch_check(ctx));
    uint32_t h = 0;
    VK_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h) {
        fprintf(stderr, "[vkzg] key load: %u of %llu bases are not points of the curve\n", h, (unsigned long long)n);
        return VKZG_ERR_ARG;
    }
    return VKZG_OK;
}

}  // namespace vk
