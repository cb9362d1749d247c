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

}  // namespace vk
