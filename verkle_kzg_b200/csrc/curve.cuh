// BN254 G1 (y^2 = x^3 + 3) group law for the device: affine inputs (64 B, the HBM table format) and
// extended-Jacobian "XYZZ" accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), which make the mixed
// addition 8M+2S and need no per-point Z for batched normalisation.
// All results leave the device as canonical affine coordinates, so the choice of internal
// representation cannot change any output byte.
#pragma once
#include "field.cuh"

namespace vk {

typedef FqParams Q;
typedef FrParams S;

struct alignas(16) affine_t {  // (0,0) encodes the point at infinity (not on the curve: 0 != 3)
    fp_t x, y;
};
struct alignas(16) xyzz_t {    // zz == 0 <=> infinity
    fp_t x, y, zz, zzz;
};

VK_HD bool affine_is_inf(const affine_t& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
VK_HD bool xyzz_is_inf(const xyzz_t& p) { return fp_is_zero(p.zz); }
VK_HD xyzz_t xyzz_inf() {
    xyzz_t r;
    r.x = fp_zero<Q>();
    r.y = fp_zero<Q>();
    r.zz = fp_zero<Q>();
    r.zzz = fp_zero<Q>();
    return r;
}
VK_HD affine_t affine_inf() {
    affine_t r;
    r.x = fp_zero<Q>();
    r.y = fp_zero<Q>();
    return r;
}
VK_HD xyzz_t xyzz_from_affine(const affine_t& p) {
    xyzz_t r;
    if (affine_is_inf(p)) return xyzz_inf();
    r.x = p.x;
    r.y = p.y;
    r.zz = fp_one<Q>();
    r.zzz = fp_one<Q>();
    return r;
}
VK_HD affine_t affine_neg(const affine_t& p) {
    affine_t r;
    r.x = p.x;
    r.y = fp_neg<Q>(p.y);  // 0 - 0 = 0 keeps infinity
    return r;
}
VK_HD xyzz_t xyzz_neg(const xyzz_t& p) {
    xyzz_t r = p;
    r.y = fp_neg<Q>(p.y);
    return r;
}

// dbl-2008-s-1 (a = 0): 6M + 3S
VK_HD xyzz_t xyzz_dbl(const xyzz_t& p) {
    if (xyzz_is_inf(p)) return p;
    fp_t U = fp_dbl<Q>(p.y);
    fp_t V = fp_sqr<Q>(U);
    fp_t W = fp_mul<Q>(U, V);
    fp_t Sx = fp_mul<Q>(p.x, V);
    fp_t X2 = fp_sqr<Q>(p.x);
    fp_t M = fp_add<Q>(fp_dbl<Q>(X2), X2);
    xyzz_t r;
    r.x = fp_sub<Q>(fp_sqr<Q>(M), fp_dbl<Q>(Sx));
    // Y3 = M (S - X3) - W Y1 as one fused pair of products (one Montgomery reduction, field.cuh: fp_mul2_lazy)
    r.y = fp_canon<Q>(fp_mul2_lazy<Q>(M, fp_sub<Q>(Sx, r.x), W, fp_neg<Q>(p.y)));
    r.zz = fp_mul<Q>(V, p.zz);
    r.zzz = fp_mul<Q>(W, p.zzz);
    return r;
}

// doubling of an affine point (ZZ = ZZZ = 1): 3M + 3S ... kept out of line, it is the rare branch of madd
__host__ __device__ __noinline__ inline xyzz_t xyzz_dbl_affine(const affine_t p) {
    xyzz_t t = xyzz_from_affine(p);
    return xyzz_dbl(t);
}

// acc += p (affine), madd-2008-s: 8M + 2S.  Complete: handles infinity, doubling and cancellation.
VK_HD void xyzz_madd(xyzz_t& acc, const affine_t& p) {
    if (affine_is_inf(p)) return;
    if (xyzz_is_inf(acc)) {
        acc = xyzz_from_affine(p);
        return;
    }
    fp_t U2 = fp_mul<Q>(p.x, acc.zz);
    fp_t S2 = fp_mul<Q>(p.y, acc.zzz);
    fp_t P = fp_sub<Q>(U2, acc.x);
    fp_t R = fp_sub<Q>(S2, acc.y);
    if (fp_is_zero(P)) {  // same x: doubling or cancellation (rare)
        if (fp_is_zero(R))
            acc = xyzz_dbl_affine(p);
        else
            acc = xyzz_inf();
        return;
    }
    fp_t PP = fp_sqr<Q>(P);
    fp_t PPP = fp_mul<Q>(P, PP);
    fp_t Qv = fp_mul<Q>(acc.x, PP);
    fp_t X3 = fp_sub<Q>(fp_sub<Q>(fp_sqr<Q>(R), PPP), fp_dbl<Q>(Qv));
    fp_t Y3 = fp_sub<Q>(fp_mul<Q>(R, fp_sub<Q>(Qv, X3)), fp_mul<Q>(acc.y, PPP));
    acc.x = X3;
    acc.y = Y3;
    acc.zz = fp_mul<Q>(acc.zz, PP);
    acc.zzz = fp_mul<Q>(acc.zzz, PPP);
}

// Same mixed addition with the ten field multiplications as CALLS of one out-of-line multiplier: the loop
// body of the accumulation kernels then is ~0.5 K instructions + one 3.5 KB multiplier instead of ~40 KB of
// inlined code, i.e. it stays inside the instruction caches (ncu showed "no instruction" stalls otherwise).
#ifdef __CUDA_ARCH__
#define VK_MUL_HOT(a, b) fp_mul_lazy_ni<Q>(a, b)  // (fully inlined was measured slower: 164 k vs 173 k proofs/s)
#define VK_MUL2_HOT(a, b, c, d) fp_mul2_lazy_ni<Q>(a, b, c, d)
#define VK_SQR_HOT(a) fp_sqr_lazy_ni<Q>(a)        // 108 instead of 136 multiply-accumulates (field.cuh: fp_sqr_lazy)
#else
#define VK_MUL_HOT(a, b) fp_mul_lazy<Q>(a, b)
#define VK_MUL2_HOT(a, b, c, d) fp_mul2_lazy<Q>(a, b, c, d)
#define VK_SQR_HOT(a) fp_sqr_lazy<Q>(a)
#endif
// The accumulator coordinates are kept in [0, 2p) ("almost Montgomery": no final conditional subtraction in the ten
// products); xyzz_canon() brings them back to [0, p) once, after the loop.  The table point is canonical.
__host__ __device__ __forceinline__ void xyzz_madd_hot(xyzz_t& acc, const affine_t& p) {
    if (affine_is_inf(p)) return;
    if (xyzz_is_inf(acc)) {
        acc = xyzz_from_affine(p);
        return;
    }
    fp_t U2 = VK_MUL_HOT(p.x, acc.zz);
    fp_t S2 = VK_MUL_HOT(p.y, acc.zzz);
    fp_t P = fp_sub_lazy<Q>(U2, acc.x);
    fp_t R = fp_sub_lazy<Q>(S2, acc.y);
    if (fp_is_zero_lazy<Q>(P)) {
        if (fp_is_zero_lazy<Q>(R))
            acc = xyzz_dbl_affine(p);
        else
            acc = xyzz_inf();
        return;
    }
    fp_t PP = VK_SQR_HOT(P);
    fp_t PPP = VK_MUL_HOT(P, PP);
    fp_t Qv = VK_MUL_HOT(acc.x, PP);
    fp_t X3 = fp_sub_lazy<Q>(fp_sub_lazy<Q>(VK_SQR_HOT(R), PPP), fp_add_lazy<Q>(Qv, Qv));
    // Y3 = R (Q - X3) - Y1 PPP as ONE fused pair of products sharing their Montgomery reduction (200 instead of 272 MACs)
    fp_t Y3 = VK_MUL2_HOT(R, fp_sub_lazy<Q>(Qv, X3), fp_neg_lazy<Q>(acc.y), PPP);
    acc.x = X3;
    acc.y = Y3;
    acc.zz = VK_MUL_HOT(acc.zz, PP);
    acc.zzz = VK_MUL_HOT(acc.zzz, PPP);
}
__host__ __device__ __forceinline__ void xyzz_canon(xyzz_t& a) {
    a.x = fp_canon<Q>(a.x);
    a.y = fp_canon<Q>(a.y);
    a.zz = fp_canon<Q>(a.zz);
    a.zzz = fp_canon<Q>(a.zzz);
}

__host__ __device__ __noinline__ inline xyzz_t xyzz_dbl_ni(const xyzz_t p) { return xyzz_dbl(p); }

// a + b, add-2008-s: 12M + 2S.  Complete.
VK_HD xyzz_t xyzz_add(const xyzz_t& a, const xyzz_t& b) {
    if (xyzz_is_inf(a)) return b;
    if (xyzz_is_inf(b)) return a;
    fp_t U1 = fp_mul<Q>(a.x, b.zz);
    fp_t U2 = fp_mul<Q>(b.x, a.zz);
    fp_t S1 = fp_mul<Q>(a.y, b.zzz);
    fp_t S2 = fp_mul<Q>(b.y, a.zzz);
    fp_t P = fp_sub<Q>(U2, U1);
    fp_t R = fp_sub<Q>(S2, S1);
    if (fp_is_zero(P)) {
        if (fp_is_zero(R)) return xyzz_dbl_ni(a);
        return xyzz_inf();
    }
    fp_t PP = fp_sqr<Q>(P);
    fp_t PPP = fp_mul<Q>(P, PP);
    fp_t Qv = fp_mul<Q>(U1, PP);
    xyzz_t r;
    r.x = fp_sub<Q>(fp_sub<Q>(fp_sqr<Q>(R), PPP), fp_dbl<Q>(Qv));
    r.y = fp_canon<Q>(fp_mul2_lazy<Q>(R, fp_sub<Q>(Qv, r.x), fp_neg<Q>(S1), PPP));  // R (Q - X3) - S1 PPP, fused pair
    r.zz = fp_mul<Q>(fp_mul<Q>(a.zz, b.zz), PP);
    r.zzz = fp_mul<Q>(fp_mul<Q>(a.zzz, b.zzz), PPP);
    return r;
}
__host__ __device__ __noinline__ inline xyzz_t xyzz_add_ni(const xyzz_t a, const xyzz_t b) { return xyzz_add(a, b); }

// Normalise with a caller-supplied inverse of zzz:  1/zz = (1/zzz)^2 * zz^2  (zz^3 = zzz^2)
VK_HD affine_t xyzz_to_affine_with_inv(const xyzz_t& p, const fp_t& zzz_inv) {
    if (xyzz_is_inf(p)) return affine_inf();
    fp_t t = fp_mul<Q>(zzz_inv, p.zz);
    fp_t zz_inv = fp_sqr<Q>(t);
    affine_t r;
    r.x = fp_mul<Q>(p.x, zz_inv);
    r.y = fp_mul<Q>(p.y, zzz_inv);
    return r;
}
__host__ __device__ inline affine_t xyzz_to_affine(const xyzz_t& p) {
    if (xyzz_is_inf(p)) return affine_inf();
    return xyzz_to_affine_with_inv(p, fp_inv<Q>(p.zzz));
}

// k * p for a small unsigned k (table seeding, bucket-range offsets); double-and-add MSB first
__host__ __device__ inline xyzz_t xyzz_mul_u32(const xyzz_t& p, uint32_t k) {
    xyzz_t acc = xyzz_inf();
    for (int b = 31; b >= 0; --b) {
        acc = xyzz_dbl_ni(acc);
        if ((k >> b) & 1) acc = xyzz_add_ni(acc, p);
    }
    return acc;
}

// ark-serialize compressed encoding of an affine point -> 32 bytes (as 8 LE words):
//   canonical x, bit 255 set if y > (p-1)/2, bit 254 set (x = 0) for infinity.
VK_HD void affine_compress(const affine_t& p, uint32_t out[8]) {
    if (affine_is_inf(p)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = 0;
        out[7] = 0x40000000u;
        return;
    }
    fp_t xc = fp_from_mont<Q>(p.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = xc.l[i];
    if (fp_is_lexicographically_largest<Q>(p.y)) out[7] |= 0x80000000u;
}

// ---------------------------------------------------------------------------------------------
// Scalar recoding: canonical 256-bit scalar -> signed c-bit digits d_w in [-(2^(c-1)-1), 2^(c-1)],
// sum_w d_w 2^(c w) = k, for W = ceil(255 / c) windows (the top window never overflows for k < 2^254).
// ---------------------------------------------------------------------------------------------
VK_HD int num_windows(int c) { return (255 + c - 1) / c; }

VK_HD uint32_t scalar_bits(const uint32_t k[8], int pos, int c) {  // c <= 24 bits starting at bit pos
    int limb = pos >> 5, sh = pos & 31;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i == limb) v |= k[i];
        if (i == limb + 1) v |= (uint64_t)k[i] << 32;
    }
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}


// d[w] for w < W; d[w] in [-(2^(c-1)-1), 2^(c-1)]
__host__ __device__ inline void recode_signed(const uint32_t k[8], int c, int W, int32_t* d) {
    uint32_t carry = 0;
    for (int w = 0; w < W; ++w) {
        uint32_t v = scalar_bits(k, w * c, c) + carry;
        if (v > (1u << (c - 1))) {
            d[w] = (int32_t)v - (int32_t)(1u << c);
            carry = 1;
        } else {
            d[w] = (int32_t)v;
            carry = 0;
        }
    }
}

}  // namespace vk
