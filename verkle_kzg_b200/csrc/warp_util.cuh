// Warp-level helpers shared by the Fr-side kernels (ipa.cu, poly.cu, multiproof.cu).
#pragma once
#include "hash.cuh"

namespace vk {

__device__ __forceinline__ fp_t shfl_fp(const fp_t& v, int src) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
__device__ __forceinline__ fp_t shfl_up_fp(const fp_t& v, int d) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
__device__ __forceinline__ fp_t shfl_down_fp(const fp_t& v, int d) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}
__device__ __forceinline__ fp_t shfl_xor_fpw(const fp_t& v, int m) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], m);
    return r;
}
__device__ __forceinline__ fp_t warp_sum_frw(fp_t v) {
#pragma unroll 1
    for (int m = 16; m > 0; m >>= 1) v = fp_add<S>(v, shfl_xor_fpw(v, m));
    return v;
}

// Every lane passes the (non-zero) product t of its own denominators; returns 1 / t per lane with ONE inversion per warp:
// prefix and suffix products over the lanes by shuffle scans, then every lane inverts the SAME grand total (uniform
// control flow — the binary inversion is several times slower when 32 lanes run it on 32 different values).
// All 32 lanes must call it.
template <class F>
static __device__ __noinline__ fp_t warp_inverse_of_lane_products_t(const fp_t t) {
    const int lane = threadIdx.x & 31;
    fp_t P = t, X = t;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        fp_t up = shfl_up_fp(P, d), dn = shfl_down_fp(X, d);
        if (lane >= d) P = fp_mul_ni<F>(P, up);
        if (lane + d < 32) X = fp_mul_ni<F>(X, dn);
    }
    fp_t inv = fp_inv<F>(shfl_fp(P, 31));
    fp_t left = shfl_up_fp(P, 1), right = shfl_down_fp(X, 1);
    if (lane > 0) inv = fp_mul_ni<F>(inv, left);
    if (lane < 31) inv = fp_mul_ni<F>(inv, right);
    return inv;
}
static __device__ __forceinline__ fp_t warp_inverse_of_lane_products(const fp_t t) { return warp_inverse_of_lane_products_t<S>(t); }

__device__ __forceinline__ xyzz_t shfl_xor_point(const xyzz_t& v, int mask) {
    xyzz_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_xor_sync(0xffffffffu, v.x.l[i], mask);
        r.y.l[i] = __shfl_xor_sync(0xffffffffu, v.y.l[i], mask);
        r.zz.l[i] = __shfl_xor_sync(0xffffffffu, v.zz.l[i], mask);
        r.zzz.l[i] = __shfl_xor_sync(0xffffffffu, v.zzz.l[i], mask);
    }
    return r;
}
__device__ __forceinline__ fp_t fp_sel(bool c, const fp_t& a, const fp_t& b) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}

// One level of a shuffle-tree fold: returns mine + (the accumulator of lane ^ off) on BOTH lanes of the pair.
// xyzz_add costs 14 products; the two lanes of a pair would compute the same 14 redundantly, so they split them
// instead — 2 + 3 + 2 products per lane with the intermediate values exchanged by shuffles — which halves the
// dependent multiplier chain of a fold level (what a latency-bound single proof waits for) and the redundant pipe
// work of the big batches.  Inputs canonical; all 32 lanes must call it (full-mask shuffles; the special cases —
// an identity operand, equal or opposite points — are resolved after the last shuffle).
static __device__ __noinline__ xyzz_t xyzz_add_pair(const xyzz_t mine, int off) {
    const bool A = ((threadIdx.x & 31) & off) == 0;  // lane A holds operand a, lane B operand b; result = a + b
    const xyzz_t other = shfl_xor_point(mine, off);
    // U1 = a.x b.zz, S1 = a.y b.zzz on A;  U2 = b.x a.zz, S2 = b.y a.zzz on B
    fp_t t1 = fp_mul_ni<Q>(mine.x, other.zz), t2 = fp_mul_ni<Q>(mine.y, other.zzz);
    fp_t o1 = shfl_xor_fpw(t1, off), o2 = shfl_xor_fpw(t2, off);
    const fp_t U1 = fp_sel(A, t1, o1), U2 = fp_sel(A, o1, t1), S1 = fp_sel(A, t2, o2), S2 = fp_sel(A, o2, t2);
    const fp_t P = fp_sub<Q>(U2, U1), R = fp_sub<Q>(S2, S1);
    // A: PP = P^2, PPP = P PP, Q = U1 PP;   B: RR = R^2, ZZ12 = a.zz b.zz, ZZZ12 = a.zzz b.zzz
    const fp_t sq = fp_sel(A, P, R);
    fp_t v1 = fp_mul_ni<Q>(sq, sq);
    fp_t v2 = fp_mul_ni<Q>(fp_sel(A, P, mine.zz), fp_sel(A, v1, other.zz));
    fp_t v3 = fp_mul_ni<Q>(fp_sel(A, U1, mine.zzz), fp_sel(A, v1, other.zzz));
    o1 = shfl_xor_fpw(v1, off);
    o2 = shfl_xor_fpw(v2, off);
    fp_t o3 = shfl_xor_fpw(v3, off);
    const fp_t PP = fp_sel(A, v1, o1), RR = fp_sel(A, o1, v1), PPP = fp_sel(A, v2, o2), ZZ12 = fp_sel(A, o2, v2);
    const fp_t Qv = fp_sel(A, v3, o3), ZZZ12 = fp_sel(A, o3, v3);
    xyzz_t r;
    r.x = fp_sub<Q>(fp_sub<Q>(RR, PPP), fp_dbl<Q>(Qv));
    // A: Y3 = R (Q - X3) - S1 PPP;   B: ZZ3 = ZZ12 PP, ZZZ3 = ZZZ12 PPP
    fp_t w1 = fp_mul_ni<Q>(fp_sel(A, R, ZZ12), fp_sel(A, fp_sub<Q>(Qv, r.x), PP));
    fp_t w2 = fp_mul_ni<Q>(fp_sel(A, S1, ZZZ12), PPP);
    o1 = shfl_xor_fpw(w1, off);
    o2 = shfl_xor_fpw(w2, off);
    r.y = fp_sub<Q>(fp_sel(A, w1, o1), fp_sel(A, w2, o2));
    r.zz = fp_sel(A, o1, w1);
    r.zzz = fp_sel(A, o2, w2);
    if (xyzz_is_inf(mine)) return other;
    if (xyzz_is_inf(other)) return mine;
    if (fp_is_zero(P)) return A ? xyzz_add_ni(mine, other) : xyzz_add_ni(other, mine);  // doubling / opposite points
    return r;
}

// ---------------------------------------------------------------------------------------------------
// Variable-base scalar multiplication k P (the verifier's C, L_r, R_r and the multiproof's commitments): fixed 4-bit
// windows, MSB first, over the table P, 2P, ..., 15P — 252 doublings + <= 64 additions instead of the 254 + ~127 of
// double-and-add, and no data-dependent branch per bit.  One thread per point: the throughput form, for batches that fill
// the GPU (a handful of points use var_mul_quad below).
// ---------------------------------------------------------------------------------------------------
static __device__ __noinline__ xyzz_t var_mul_windowed(const affine_t P, const fp_t k_canon) {
    if (affine_is_inf(P)) return xyzz_inf();
    xyzz_t T[15];
    T[0] = xyzz_from_affine(P);
#pragma unroll 1
    for (int d = 2; d <= 15; ++d) T[d - 1] = (d & 1) ? xyzz_add_ni(T[d - 2], T[0]) : xyzz_dbl_ni(T[d / 2 - 1]);
    xyzz_t acc = xyzz_inf();
#pragma unroll 1
    for (int w = 63; w >= 0; --w) {
        if (!xyzz_is_inf(acc)) {
#pragma unroll 1
            for (int j = 0; j < 4; ++j) acc = xyzz_dbl_ni(acc);
        }
        uint32_t limb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i == (w >> 3)) limb = k_canon.l[i];
        const uint32_t d = (limb >> (4 * (w & 7))) & 15;
        if (d) acc = xyzz_add_ni(acc, T[d - 1]);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------------
// Variable-base scalar multiplication in latency mode: FOUR lanes per point.
// k P needs ~252 dependent doublings whatever the algorithm, and a lone thread is a chain of dependent calls of the
// out-of-line multiplier (~980 cycles each, tools/coop_probe.cu; inlining the products instead makes 30-50 KB of
// straight-line code per point operation, which a lone warp cannot fetch any faster).  The products of one doubling
// have dependency depth 3 and of one addition depth 4, so four lanes holding the same accumulator compute one product
// each per level and exchange the results inside their quad (width-4 shuffles): measured 5.2 k cycles per doubling and
// 9.8 k per addition against 8.8 k / 13.7 k.  All 32 lanes of a warp must call these (full-mask shuffles); special
// cases are resolved after the last shuffle.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ fp_t quad_get(const fp_t& v, int q) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], q, 4);
    return r;
}
// fp_sel over four choices by the lane's position in its quad
__device__ __forceinline__ fp_t quad_pick(int q, const fp_t& a0, const fp_t& a1, const fp_t& a2, const fp_t& a3) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = q == 0 ? a0.l[i] : q == 1 ? a1.l[i] : q == 2 ? a2.l[i] : a3.l[i];
    return r;
}

// 2 * acc (dbl-2008-s-1, a = 0): levels {V = U^2, XX = X^2}, {W = U V, S = X V, MM = M^2, ZZ3 = V ZZ}, {Y3a, Y3b, ZZZ3}
static __device__ __noinline__ xyzz_t xyzz_dbl_quad(const xyzz_t p) {
    const int q = threadIdx.x & 3;
    const fp_t U = fp_dbl<Q>(p.y);
    const fp_t sq = (q & 1) ? p.x : U;
    const fp_t l1 = fp_mul_ni<Q>(sq, sq);
    const fp_t V = quad_get(l1, 0), XX = quad_get(l1, 1);
    const fp_t M = fp_add<Q>(fp_dbl<Q>(XX), XX);
    const fp_t l2 = fp_mul_ni<Q>(quad_pick(q, U, p.x, M, V), quad_pick(q, V, V, M, p.zz));
    const fp_t W = quad_get(l2, 0), Sx = quad_get(l2, 1), MM = quad_get(l2, 2);
    xyzz_t r;
    r.zz = quad_get(l2, 3);
    r.x = fp_sub<Q>(MM, fp_dbl<Q>(Sx));
    const fp_t l3 = fp_mul_ni<Q>(quad_pick(q, M, W, W, W), quad_pick(q, fp_sub<Q>(Sx, r.x), p.y, p.zzz, p.zzz));
    r.y = fp_sub<Q>(quad_get(l3, 0), quad_get(l3, 1));
    r.zzz = quad_get(l3, 2);
    return xyzz_is_inf(p) ? p : r;
}

// acc + P (madd-2008-s): levels {U2 = x2 ZZ, S2 = y2 ZZZ}, {PP = P^2, RR = R^2}, {PPP = P PP, Q = X PP, ZZ3 = ZZ PP},
// {Y3a = R (Q - X3), Y3b = Y PPP, ZZZ3 = ZZZ PPP}
static __device__ __noinline__ xyzz_t xyzz_madd_quad(const xyzz_t acc, const affine_t pt) {
    const int q = threadIdx.x & 3;
    const fp_t l1 = fp_mul_ni<Q>((q & 1) ? pt.y : pt.x, (q & 1) ? acc.zzz : acc.zz);
    const fp_t U2 = quad_get(l1, 0), S2 = quad_get(l1, 1);
    const fp_t P = fp_sub<Q>(U2, acc.x), R = fp_sub<Q>(S2, acc.y);
    const fp_t sq = (q & 1) ? R : P;
    const fp_t l2 = fp_mul_ni<Q>(sq, sq);
    const fp_t PP = quad_get(l2, 0), RR = quad_get(l2, 1);
    const fp_t l3 = fp_mul_ni<Q>(quad_pick(q, P, acc.x, acc.zz, acc.zz), PP);
    const fp_t PPP = quad_get(l3, 0), Qv = quad_get(l3, 1);
    xyzz_t r;
    r.zz = quad_get(l3, 2);
    r.x = fp_sub<Q>(fp_sub<Q>(RR, PPP), fp_dbl<Q>(Qv));
    const fp_t l4 = fp_mul_ni<Q>(quad_pick(q, R, acc.y, acc.zzz, acc.zzz), quad_pick(q, fp_sub<Q>(Qv, r.x), PPP, PPP, PPP));
    r.y = fp_sub<Q>(quad_get(l4, 0), quad_get(l4, 1));
    r.zzz = quad_get(l4, 2);
    if (affine_is_inf(pt)) return acc;
    if (xyzz_is_inf(acc)) return xyzz_from_affine(pt);
    if (fp_is_zero(P)) return fp_is_zero(R) ? xyzz_dbl_affine(pt) : xyzz_inf();
    return r;
}

// a + b, both XYZZ (add-2008-s): levels {U1, U2, S1, S2}, {PP, RR, ZZ12 = ZZ1 ZZ2, ZZZ12 = ZZZ1 ZZZ2},
// {PPP, Q = U1 PP, ZZ3 = ZZ12 PP}, {Y3a = R (Q - X3), Y3b = S1 PPP, ZZZ3 = ZZZ12 PPP}
static __device__ __noinline__ xyzz_t xyzz_add_quad(const xyzz_t a, const xyzz_t b) {
    const int q = threadIdx.x & 3;
    const fp_t l1 = fp_mul_ni<Q>(quad_pick(q, a.x, b.x, a.y, b.y), quad_pick(q, b.zz, a.zz, b.zzz, a.zzz));
    const fp_t U1 = quad_get(l1, 0), U2 = quad_get(l1, 1), S1 = quad_get(l1, 2), S2 = quad_get(l1, 3);
    const fp_t P = fp_sub<Q>(U2, U1), R = fp_sub<Q>(S2, S1);
    const fp_t l2 = fp_mul_ni<Q>(quad_pick(q, P, R, a.zz, a.zzz), quad_pick(q, P, R, b.zz, b.zzz));
    const fp_t PP = quad_get(l2, 0), RR = quad_get(l2, 1), ZZ12 = quad_get(l2, 2), ZZZ12 = quad_get(l2, 3);
    const fp_t l3 = fp_mul_ni<Q>(quad_pick(q, P, U1, ZZ12, ZZ12), PP);
    const fp_t PPP = quad_get(l3, 0), Qv = quad_get(l3, 1);
    xyzz_t r;
    r.zz = quad_get(l3, 2);
    r.x = fp_sub<Q>(fp_sub<Q>(RR, PPP), fp_dbl<Q>(Qv));
    const fp_t l4 = fp_mul_ni<Q>(quad_pick(q, R, S1, ZZZ12, ZZZ12), quad_pick(q, fp_sub<Q>(Qv, r.x), PPP, PPP, PPP));
    r.y = fp_sub<Q>(quad_get(l4, 0), quad_get(l4, 1));
    r.zzz = quad_get(l4, 2);
    if (xyzz_is_inf(a)) return b;
    if (xyzz_is_inf(b)) return a;
    if (fp_is_zero(P)) return xyzz_add_ni(a, b);  // doubling / opposite points
    return r;
}

// k * P for the quad's point (all four lanes pass the same P and the same canonical scalar k < 2^254): fixed 4-bit
// windows, MSB first, over the table P, 2P, ..., 15P (7 doublings + 7 mixed additions to build) — per window four
// doublings (3 levels each) and one addition (4 levels): ~1070 multiplier latencies instead of the ~4800 of a lone thread.
static __device__ __noinline__ xyzz_t var_mul_quad(const affine_t P, const fp_t k_canon) {
    xyzz_t T[15];
    T[0] = xyzz_from_affine(P);
    if (affine_is_inf(P)) T[0] = xyzz_inf();
#pragma unroll 1
    for (int d = 2; d <= 15; ++d) T[d - 1] = (d & 1) ? xyzz_madd_quad(T[d - 2], P) : xyzz_dbl_quad(T[d / 2 - 1]);
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
        if (__any_sync(0xffffffffu, d != 0)) {  // (warp-uniform branch: the shuffles inside need every lane)
            xyzz_t t = xyzz_add_quad(acc, T[d ? d - 1 : 0]);
            if (d) acc = t;
        }
    }
    return acc;
}

}  // namespace vk
