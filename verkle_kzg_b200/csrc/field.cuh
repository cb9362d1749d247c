// BN254 Fr / Fq arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256
// (bit-identical in memory to arkworks' 4 x u64 MontBackend).
//
// The multiplier is an operand-scanning Montgomery product that keeps the running total split over
// two limb arrays whose 64-bit (lo,hi) pairs sit at even / odd 32-bit columns, so every partial
// product a[j]*b[i] is ONE mad.lo.cc/madc.hi.cc pair on an aligned register pair (ptxas fuses the
// pair into IMAD.WIDE.U32[.X]) and a row consists of two independent carry chains of four such
// pairs.  The 32-bit right shift of the running total after each reduction step is a swap of the
// two arrays' roles, not a data move.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define VK_HD __host__ __device__ __forceinline__

namespace vk {

struct alignas(16) fp_t {
    uint32_t l[8];
};

struct FrParams {
    static constexpr uint32_t INV = 0xefffffffu;
    VK_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    VK_HD static constexpr uint32_t r1(int i) {
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    VK_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
    VK_HD static constexpr uint32_t half(int i) {  // (p-1)/2
        constexpr uint32_t v[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
        return v[i];
    }
    VK_HD static constexpr uint32_t r3(int i) {  // R^3 mod p
        constexpr uint32_t v[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
        return v[i];
    }
};

struct FqParams {
    static constexpr uint32_t INV = 0xe4866389u;
    VK_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    VK_HD static constexpr uint32_t r1(int i) {
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    VK_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
    VK_HD static constexpr uint32_t half(int i) {
        constexpr uint32_t v[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
        return v[i];
    }
    VK_HD static constexpr uint32_t r3(int i) {  // R^3 mod p
        constexpr uint32_t v[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u};
        return v[i];
    }
};

// ---------------------------------------------------------------------------------------------
// carry-chain building blocks.  Device: one PTX asm block per chain (self-contained w.r.t. the CC
// flag).  Host: the same semantics emulated with 64-bit arithmetic, so that every function in this
// header can be unit-tested on a CPU (tests/host/) with the identical control flow.
// ---------------------------------------------------------------------------------------------
#ifndef __CUDA_ARCH__
VK_HD void host_mad_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t addlo, uint32_t addhi, uint32_t& carry) {
    // (carry, hi, lo) = a*b + (addhi:addlo) + carry   with carries propagated like mad.lo.cc / madc.hi.cc
    uint64_t prod = (uint64_t)a * b;
    uint64_t l = (uint64_t)(uint32_t)prod + addlo + carry;
    uint64_t h = (prod >> 32) + addhi + (l >> 32);
    lo = (uint32_t)l;
    hi = (uint32_t)h;
    carry = (uint32_t)(h >> 32);
}
#endif

// acc[0..7] += {a0,a2,a4,a6} * b at 64-bit column pairs (0,1),(2,3),(4,5),(6,7); top += carry-out
VK_HD void mad_row4(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
    uint32_t c = 0;
    host_mad_pair(acc[0], acc[1], a0, b, acc[0], acc[1], c);
    host_mad_pair(acc[2], acc[3], a2, b, acc[2], acc[3], c);
    host_mad_pair(acc[4], acc[5], a4, b, acc[4], acc[5], c);
    host_mad_pair(acc[6], acc[7], a6, b, acc[6], acc[7], c);
    top += c;
#endif
}

// Same, without a carry-out (caller guarantees none can occur).
VK_HD void mad_row4_nc(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
#else
    uint32_t c = 0;
    host_mad_pair(acc[0], acc[1], a0, b, acc[0], acc[1], c);
    host_mad_pair(acc[2], acc[3], a2, b, acc[2], acc[3], c);
    host_mad_pair(acc[4], acc[5], a4, b, acc[4], acc[5], c);
    host_mad_pair(acc[6], acc[7], a6, b, acc[6], acc[7], c);
#endif
}

// Row entry after a role swap:  x0 += e1 (column 0), the carry continues into the column-1 chain
//   y[k,k+1] = e[k+2,k+3] + a_odd * b    for k = 0,2,4 ;   y[6,7] = a7 * b + carry
// where e[] is the array whose column 0 was just cleared by the reduction step.
VK_HD void shift_mad_row4(uint32_t& x0, uint32_t* y, const uint32_t* e, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7,
                          uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %1, %16, %20, %10;\n\t"
        "madc.hi.cc.u32 %2, %16, %20, %11;\n\t"
        "madc.lo.cc.u32 %3, %17, %20, %12;\n\t"
        "madc.hi.cc.u32 %4, %17, %20, %13;\n\t"
        "madc.lo.cc.u32 %5, %18, %20, %14;\n\t"
        "madc.hi.cc.u32 %6, %18, %20, %15;\n\t"
        "madc.lo.cc.u32 %7, %19, %20, 0;\n\t"
        "madc.hi.u32 %8, %19, %20, 0;"
        : "+r"(x0), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
#else
    uint64_t s = (uint64_t)x0 + e[1];
    x0 = (uint32_t)s;
    uint32_t c = (uint32_t)(s >> 32);
    host_mad_pair(y[0], y[1], a1, b, e[2], e[3], c);
    host_mad_pair(y[2], y[3], a3, b, e[4], e[5], c);
    host_mad_pair(y[4], y[5], a5, b, e[6], e[7], c);
    host_mad_pair(y[6], y[7], a7, b, 0, 0, c);
#endif
}

// r[0..7] = a[0..7] + b[0..7], returns the carry-out
VK_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t carry;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(carry)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]),
          "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    carry = (uint32_t)c;
#endif
    return carry;
}

// r[0..7] = a[0..7] - b[0..7], returns 0xffffffff on borrow-out else 0
VK_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t borrow;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]),
          "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint32_t bw = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - bw;
        r[i] = (uint32_t)d;
        bw = (uint32_t)(d >> 32) & 1;
    }
    borrow = bw ? 0xffffffffu : 0u;
#endif
    return borrow;
}

// (carry : r) >= p ?  r -= p   — r < 2p, carry is the 9th limb (0 or 1)
template <class P>
VK_HD void fp_cond_sub_p(uint32_t* r, uint32_t carry) {
    uint32_t d[8], pl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pl[k] = P::p(k);
    uint32_t borrow = sub8(d, r, pl);
    // keep r only if the subtraction borrowed AND there was no 9th-limb carry
    bool keep = (borrow != 0) && (carry == 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = keep ? r[k] : d[k];
}

template <class P>
VK_HD void reduce_step(uint32_t* x, uint32_t* y) {
    // m = x[0] * (-p^-1) ;  total += m * p  -> x[0] becomes 0
    uint32_t m = x[0] * P::INV;
    mad_row4_nc(y, P::p(1), P::p(3), P::p(5), P::p(7), m);
    mad_row4(x, y[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
}

// r = a * b * R^-1 mod p, fully reduced to [0, p).  Requires a < p; b may be ANY 256-bit value (the running
// total stays below a + p < 2p, which is what the no-carry-out chains rely on).
template <class P>
VK_HD fp_t fp_mul(const fp_t& a, const fp_t& b) {
    uint32_t u[8], v[8];
    // row 0: u (column-0 aligned) = even limbs of a times b0, v (column-1 aligned) = odd limbs
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        uint64_t t0 = (uint64_t)a.l[j] * b.l[0];
        uint64_t t1 = (uint64_t)a.l[j + 1] * b.l[0];
        u[j] = (uint32_t)t0;
        u[j + 1] = (uint32_t)(t0 >> 32);
        v[j] = (uint32_t)t1;
        v[j + 1] = (uint32_t)(t1 >> 32);
    }
    reduce_step<P>(u, v);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        // roles swap: v becomes the column-0 array, a fresh column-1 array y is built from u >> 64
        {
            uint32_t y[8];
            shift_mad_row4(v[0], y, u, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            mad_row4(v, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            reduce_step<P>(v, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = y[k];
        }
        if (i + 1 < 8) {
            uint32_t y[8];
            shift_mad_row4(u[0], y, v, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
            mad_row4(u, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
            reduce_step<P>(u, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = y[k];
        }
    }
    // now v is the column-0 array with v[0] == 0 and u the column-1 array:  result = (v >> 32) + u  (< 2p)
    uint32_t vs[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) vs[k] = v[k + 1];
    vs[7] = 0;
    fp_t r;
    add8(r.l, u, vs);
    fp_cond_sub_p<P>(r.l, 0);
    return r;
}

// "Almost Montgomery" product for the hot loops: inputs in [0, 2p), output in [0, 2p), NO final conditional
// subtraction (R = 2^256 > 4p: (ab + mp)/R < 4p^2/R + p < 2p; the running total stays below 3p < 2^256).
template <class P>
VK_HD fp_t fp_mul_lazy(const fp_t& a, const fp_t& b) {
    uint32_t u[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        uint64_t t0 = (uint64_t)a.l[j] * b.l[0];
        uint64_t t1 = (uint64_t)a.l[j + 1] * b.l[0];
        u[j] = (uint32_t)t0;
        u[j + 1] = (uint32_t)(t0 >> 32);
        v[j] = (uint32_t)t1;
        v[j + 1] = (uint32_t)(t1 >> 32);
    }
    reduce_step<P>(u, v);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        {
            uint32_t y[8];
            shift_mad_row4(v[0], y, u, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            mad_row4(v, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            reduce_step<P>(v, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = y[k];
        }
        if (i + 1 < 8) {
            uint32_t y[8];
            shift_mad_row4(u[0], y, v, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
            mad_row4(u, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
            reduce_step<P>(u, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = y[k];
        }
    }
    uint32_t vs[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) vs[k] = v[k + 1];
    vs[7] = 0;
    fp_t r;
    add8(r.l, u, vs);
    return r;
}
template <class P>
__host__ __device__ __noinline__ fp_t fp_mul_lazy_ni(const fp_t a, const fp_t b) {
    return fp_mul_lazy<P>(a, b);
}
// Fused pair of products for the hot loops:  (a b + c d) R^-1 mod p  in ONE interleaved pass — the two row products of a
// step share its reduction row, 8 x (8 + 8 + 8 + 1) = 200 multiply-accumulates instead of the 272 of two separate products
// (Y3 = R (Q - X3) + (-Y1) PPP of the mixed addition).  Inputs in [0, 2p]; the running total stays below 5p (1 + 2^-32)
// < 2^256 after every shift and below 2^288 inside a row, so neither array can carry out of its top limb beyond what
// `top` catches; the result (a b + c d + M p) / R < 8 p^2 / R + p < 2.51 p is brought back to [0, 2p) by one conditional
// subtraction of 2p.
template <class P>
VK_HD fp_t fp_mul2_lazy(const fp_t& a, const fp_t& b, const fp_t& c, const fp_t& d) {
    uint32_t u[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        uint64_t t0 = (uint64_t)a.l[j] * b.l[0];
        uint64_t t1 = (uint64_t)a.l[j + 1] * b.l[0];
        u[j] = (uint32_t)t0;
        u[j + 1] = (uint32_t)(t0 >> 32);
        v[j] = (uint32_t)t1;
        v[j + 1] = (uint32_t)(t1 >> 32);
    }
    mad_row4_nc(v, c.l[1], c.l[3], c.l[5], c.l[7], d.l[0]);
    mad_row4(u, v[7], c.l[0], c.l[2], c.l[4], c.l[6], d.l[0]);
    reduce_step<P>(u, v);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        {
            uint32_t y[8];
            shift_mad_row4(v[0], y, u, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            mad_row4(v, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            mad_row4_nc(y, c.l[1], c.l[3], c.l[5], c.l[7], d.l[i]);
            mad_row4(v, y[7], c.l[0], c.l[2], c.l[4], c.l[6], d.l[i]);
            reduce_step<P>(v, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = y[k];
        }
        if (i + 1 < 8) {
            uint32_t y[8];
            shift_mad_row4(u[0], y, v, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
            mad_row4(u, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
            mad_row4_nc(y, c.l[1], c.l[3], c.l[5], c.l[7], d.l[i + 1]);
            mad_row4(u, y[7], c.l[0], c.l[2], c.l[4], c.l[6], d.l[i + 1]);
            reduce_step<P>(u, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = y[k];
        }
    }
    uint32_t vs[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) vs[k] = v[k + 1];
    vs[7] = 0;
    fp_t r;
    add8(r.l, u, vs);
    uint32_t dd[8], p2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) p2[k] = (P::p(k) << 1) | (k ? P::p(k - 1) >> 31 : 0);
    uint32_t borrow = sub8(dd, r.l, p2);
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = borrow ? r.l[k] : dd[k];
    return r;
}
template <class P>
__host__ __device__ __noinline__ fp_t fp_mul2_lazy_ni(const fp_t a, const fp_t b, const fp_t c, const fp_t d) {
    return fp_mul2_lazy<P>(a, b, c, d);
}
// ---------------------------------------------------------------------------------------------
// Dedicated squaring for the hot loops.  a^2 = sum_i a_i * V_i * 2^(32 i) with the row vectors
//     V_i = [0 .. 0, a_i, d_(i+1) & ~1, d_(i+2), .. , d_7]      (d = 2a as an 8-limb number; a <= 2p < 2^255)
// i.e. row i carries its diagonal term and the DOUBLED off-diagonal terms to its right ((2a) >> 32(i+1) equals
// 2 (a >> 32(i+1)) + bit 31 of a_i: clearing bit 0 of d_(i+1) removes the stray bit).  Row i has 8 - i products instead
// of 8 — 36 instead of 64 for the whole square — and is interleaved with the same reduction rows as fp_mul_lazy:
// 36 + 64 + 8 = 108 multiply-accumulates instead of 136.  The running total obeys the bound of the fused pair
// (t_i < t_(i-1) / 2^32 + 5p, V_i <= 4p), the result (a^2 + M p) / R < 4p^2 / R + p < 2p needs no subtraction.
// ---------------------------------------------------------------------------------------------
// pairs K0..3 of mad_row4 (the multiplicands of the pairs below K0 are zero)
template <int K0>
VK_HD void mad_row_from(uint32_t* acc, uint32_t& top, const uint32_t* m, uint32_t b) {
#ifdef __CUDA_ARCH__
    if constexpr (K0 == 0) {
        mad_row4(acc, top, m[0], m[1], m[2], m[3], b);
    } else if constexpr (K0 == 1) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, %6, 0;"
            : "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(b));
    } else if constexpr (K0 == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(m[2]), "r"(m[3]), "r"(b));
    } else if constexpr (K0 == 3) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(m[3]), "r"(b));
    }
#else
    uint32_t c = 0;
    for (int k = K0; k < 4; ++k) host_mad_pair(acc[2 * k], acc[2 * k + 1], m[k], b, acc[2 * k], acc[2 * k + 1], c);
    top += c;
#endif
}

// shift_mad_row4 whose pairs below K0 have zero multiplicands: those pairs only pass e[] (and the carry) on
template <int K0>
VK_HD void shift_mad_row_from(uint32_t& x0, uint32_t* y, const uint32_t* e, const uint32_t* m, uint32_t b) {
#ifdef __CUDA_ARCH__
    if constexpr (K0 == 0) {
        shift_mad_row4(x0, y, e, m[0], m[1], m[2], m[3], b);
    } else if constexpr (K0 == 1) {
        asm("add.cc.u32 %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %10, 0;\n\t"
            "addc.cc.u32 %2, %11, 0;\n\t"
            "madc.lo.cc.u32 %3, %16, %19, %12;\n\t"
            "madc.hi.cc.u32 %4, %16, %19, %13;\n\t"
            "madc.lo.cc.u32 %5, %17, %19, %14;\n\t"
            "madc.hi.cc.u32 %6, %17, %19, %15;\n\t"
            "madc.lo.cc.u32 %7, %18, %19, 0;\n\t"
            "madc.hi.u32 %8, %18, %19, 0;"
            : "+r"(x0), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(b));
    } else if constexpr (K0 == 2) {
        asm("add.cc.u32 %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %10, 0;\n\t"
            "addc.cc.u32 %2, %11, 0;\n\t"
            "addc.cc.u32 %3, %12, 0;\n\t"
            "addc.cc.u32 %4, %13, 0;\n\t"
            "madc.lo.cc.u32 %5, %16, %18, %14;\n\t"
            "madc.hi.cc.u32 %6, %16, %18, %15;\n\t"
            "madc.lo.cc.u32 %7, %17, %18, 0;\n\t"
            "madc.hi.u32 %8, %17, %18, 0;"
            : "+r"(x0), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(m[2]), "r"(m[3]), "r"(b));
    } else {
        static_assert(K0 == 3, "row 7 is the last one");
        asm("add.cc.u32 %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %10, 0;\n\t"
            "addc.cc.u32 %2, %11, 0;\n\t"
            "addc.cc.u32 %3, %12, 0;\n\t"
            "addc.cc.u32 %4, %13, 0;\n\t"
            "addc.cc.u32 %5, %14, 0;\n\t"
            "addc.cc.u32 %6, %15, 0;\n\t"
            "madc.lo.cc.u32 %7, %16, %17, 0;\n\t"
            "madc.hi.u32 %8, %16, %17, 0;"
            : "+r"(x0), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(m[3]), "r"(b));
    }
#else
    uint64_t s = (uint64_t)x0 + e[1];
    x0 = (uint32_t)s;
    uint32_t c = (uint32_t)(s >> 32);
    for (int k = 0; k < 4; ++k) {
        const uint32_t lo = k < 3 ? e[2 * k + 2] : 0u, hi = k < 3 ? e[2 * k + 3] : 0u;
        host_mad_pair(y[2 * k], y[2 * k + 1], k < K0 ? 0u : m[k], b, lo, hi, c);
    }
#endif
}

// limb j of the row vector V_I (see above); zero below the diagonal
template <int I>
VK_HD uint32_t sqr_row_limb(const uint32_t* a, const uint32_t* d, int j) {
    return j < I ? 0u : j == I ? a[I] : j == I + 1 ? (d[j] & ~1u) : d[j];
}
// row I >= 1 of the square: x = the array that becomes column-0 aligned, e = the previous column-0 array, y = the new column-1 array
template <class P, int I>
VK_HD void sqr_row(uint32_t* x, const uint32_t* e, uint32_t* y, const uint32_t* a, const uint32_t* d) {
    uint32_t me[4], mo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        me[k] = sqr_row_limb<I>(a, d, 2 * k);
        mo[k] = sqr_row_limb<I>(a, d, 2 * k + 1);
    }
    shift_mad_row_from<I / 2>(x[0], y, e, mo, a[I]);
    if constexpr (I < 7) mad_row_from<(I + 1) / 2>(x, y[7], me, a[I]);
    reduce_step<P>(x, y);
}
// a^2 R^-1 mod p in [0, 2p) for a in [0, 2p]
template <class P>
VK_HD fp_t fp_sqr_lazy(const fp_t& a) {
    uint32_t d[8], u[8], v[8], y[8];
#pragma unroll
    for (int k = 7; k > 0; --k) d[k] = (a.l[k] << 1) | (a.l[k - 1] >> 31);
    d[0] = a.l[0] << 1;
    // row 0: V_0 = [a0, d1 & ~1, d2 .. d7] times a0
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        uint64_t t0 = (uint64_t)sqr_row_limb<0>(a.l, d, j) * a.l[0];
        uint64_t t1 = (uint64_t)sqr_row_limb<0>(a.l, d, j + 1) * a.l[0];
        u[j] = (uint32_t)t0;
        u[j + 1] = (uint32_t)(t0 >> 32);
        v[j] = (uint32_t)t1;
        v[j + 1] = (uint32_t)(t1 >> 32);
    }
    reduce_step<P>(u, v);
#define VK_SQR_ROWS(I, X, E)                 \
    sqr_row<P, I>(X, E, y, a.l, d);          \
    _Pragma("unroll") for (int k = 0; k < 8; ++k) E[k] = y[k];
    VK_SQR_ROWS(1, v, u)
    VK_SQR_ROWS(2, u, v)
    VK_SQR_ROWS(3, v, u)
    VK_SQR_ROWS(4, u, v)
    VK_SQR_ROWS(5, v, u)
    VK_SQR_ROWS(6, u, v)
    VK_SQR_ROWS(7, v, u)
#undef VK_SQR_ROWS
    uint32_t vs[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) vs[k] = v[k + 1];
    vs[7] = 0;
    fp_t r;
    add8(r.l, u, vs);
    return r;
}
template <class P>
__host__ __device__ __noinline__ fp_t fp_sqr_lazy_ni(const fp_t a) {
    return fp_sqr_lazy<P>(a);
}
// canonical square (a <= 2p in, [0, p) out) through the dedicated square
template <class P>
VK_HD fp_t fp_sqr(const fp_t& a) {
    fp_t r = fp_sqr_lazy<P>(a);
    fp_cond_sub_p<P>(r.l, 0);
    return r;
}
// 2p - a for a in [0, 2p): the negation of a lazy value, in (0, 2p]
template <class P>
VK_HD fp_t fp_neg_lazy(const fp_t& a) {
    uint32_t p2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) p2[k] = (P::p(k) << 1) | (k ? P::p(k - 1) >> 31 : 0);
    fp_t r;
    sub8(r.l, p2, a.l);
    return r;
}
// lazy add / sub / double on [0, 2p)
template <class P>
VK_HD fp_t fp_sub_lazy(const fp_t& a, const fp_t& b) {
    fp_t r;
    uint32_t borrow = sub8(r.l, a.l, b.l);
    uint32_t pm[8], t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pm[k] = borrow & ((P::p(k) << 1) | (k ? P::p(k - 1) >> 31 : 0));  // 2p
    add8(t, r.l, pm);
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = t[k];
    return r;
}
template <class P>
VK_HD fp_t fp_add_lazy(const fp_t& a, const fp_t& b) {
    fp_t r;
    add8(r.l, a.l, b.l);  // < 4p < 2^256: no carry out
    uint32_t d[8], p2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) p2[k] = (P::p(k) << 1) | (k ? P::p(k - 1) >> 31 : 0);
    uint32_t borrow = sub8(d, r.l, p2);
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = borrow ? r.l[k] : d[k];
    return r;
}
// [0, 2p) -> [0, p)
template <class P>
VK_HD fp_t fp_canon(const fp_t& a) {
    fp_t r = a;
    fp_cond_sub_p<P>(r.l, 0);
    return r;
}
// a == 0 (mod p) for a in [0, 2p)
template <class P>
VK_HD bool fp_is_zero_lazy(const fp_t& a) {
    uint32_t z = 0, e = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        z |= a.l[k];
        e |= a.l[k] ^ P::p(k);
    }
    return z == 0 || e == 0;
}


template <class P>
VK_HD fp_t fp_add(const fp_t& a, const fp_t& b) {
    fp_t r;
    uint32_t carry = add8(r.l, a.l, b.l);
    fp_cond_sub_p<P>(r.l, carry);
    return r;
}

template <class P>
VK_HD fp_t fp_sub(const fp_t& a, const fp_t& b) {
    fp_t r;
    uint32_t borrow = sub8(r.l, a.l, b.l);
    uint32_t pm[8], t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pm[k] = borrow & P::p(k);
    add8(t, r.l, pm);
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = t[k];
    return r;
}

template <class P>
VK_HD fp_t fp_zero() {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = 0;
    return r;
}
template <class P>
VK_HD fp_t fp_one() {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = P::r1(i);
    return r;
}
VK_HD bool fp_is_zero(const fp_t& a) {
    return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
}
VK_HD bool fp_eq(const fp_t& a, const fp_t& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ b.l[i];
    return d == 0;
}
template <class P>
VK_HD fp_t fp_neg(const fp_t& a) {
    return fp_sub<P>(fp_zero<P>(), a);
}
template <class P>
VK_HD fp_t fp_dbl(const fp_t& a) {
    return fp_add<P>(a, a);
}
// Montgomery -> canonical integer
template <class P>
VK_HD fp_t fp_from_mont(const fp_t& a) {
    fp_t one;
#pragma unroll
    for (int i = 0; i < 8; ++i) one.l[i] = (i == 0);
    return fp_mul<P>(a, one);
}
// canonical integer (< p) -> Montgomery
template <class P>
VK_HD fp_t fp_to_mont(const fp_t& a) {
    fp_t r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.l[i] = P::r2(i);
    return fp_mul<P>(a, r2);
}
template <class P>
VK_HD fp_t fp_from_u32(uint32_t x) {
    fp_t a;
#pragma unroll
    for (int i = 0; i < 8; ++i) a.l[i] = (i == 0) ? x : 0;
    return fp_to_mont<P>(a);
}
// lexicographic compare of raw limb arrays (use on canonical values): a > b
VK_HD bool limbs_gt(const uint32_t* a, const uint32_t* b) {
    bool gt = false, decided = false;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (!decided && a[i] != b[i]) {
            gt = a[i] > b[i];
            decided = true;
        }
    }
    return gt;
}
// canonical(a) > (p-1)/2   (ark-ec SWFlags::from_y_coordinate: y > -y)
template <class P>
VK_HD bool fp_is_lexicographically_largest(const fp_t& a_mont) {
    fp_t c = fp_from_mont<P>(a_mont);
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = P::half(i);
    return limbs_gt(c.l, h);
}

// Out-of-line multiply for cold paths (keeps code size down where throughput does not matter).
template <class P>
__host__ __device__ __noinline__ fp_t fp_mul_ni(const fp_t a, const fp_t b) {
    return fp_mul<P>(a, b);
}

// a^(p-2): Fermat inversion (0 -> 0).  Uniform control flow; ~256 squarings + ~130 multiplies.  Kept as the
// cross-check of the binary inversion below (tests/host).
template <class P>
__host__ __device__ __noinline__ fp_t fp_inv_fermat(const fp_t a) {
    fp_t acc = fp_one<P>();
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        const uint32_t e = P::p(k) - (k == 0 ? 2u : 0u);  // p[0] >= 2 for both moduli
#pragma unroll 1
        for (int b = 31; b >= 0; --b) {
            acc = fp_mul_ni<P>(acc, acc);
            if ((e >> b) & 1) acc = fp_mul_ni<P>(acc, a);
        }
    }
    return acc;
}

// x >>= 1 over 8 limbs (top bit filled with `top`)
VK_HD void shr1_8(uint32_t* x, uint32_t top) {
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[7] = (x[7] >> 1) | (top << 31);
}
VK_HD bool geq8(const uint32_t* a, const uint32_t* b) {  // a >= b
    uint32_t t[8];
    return sub8(t, a, b) == 0;
}

VK_HD void shl1_8(uint32_t* x) {
#pragma unroll
    for (int i = 7; i > 0; --i) x[i] = (x[i] << 1) | (x[i - 1] >> 31);
    x[0] <<= 1;
}

// x is even and non-zero: x >>= z, y <<= z, k += z for z = the number of trailing zero bits of x (31 at a time)
VK_HD void strip_zeros8(uint32_t* x, uint32_t* y, uint32_t& k) {
    while ((x[0] & 1) == 0) {
#ifdef __CUDA_ARCH__
        const uint32_t z = x[0] ? (uint32_t)__ffs((int)x[0]) - 1 : 31;  // 1..31
#else
        const uint32_t z = x[0] ? (uint32_t)__builtin_ctz(x[0]) : 31;
#endif
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = (x[i] >> z) | (x[i + 1] << (32 - z));
        x[7] >>= z;
#pragma unroll
        for (int i = 7; i > 0; --i) y[i] = (y[i] << z) | (y[i - 1] >> (32 - z));
        y[0] <<= z;
        k += z;
    }
}

// Modular inverse: Kaliski's "almost Montgomery inverse" on the Montgomery REPRESENTATIVE a~ = aR — every step is a
// shift of u or v and a shift of r or s (no modular correction inside the loop), k in [254, 508] steps — gives
// a~^-1 2^k = a^-1 R^-1 2^k; multiplying by 2^(512 - k) lands on a^-1 R, the Montgomery form of the inverse.  0 -> 0.
// ~3x fewer dependent instructions than a binary Euclid that halves its cofactors mod p, ~10x fewer than the Fermat
// ladder (fp_inv_fermat, kept as the cross-check in tests/host).  Data-dependent control flow: where many values are
// inverted, warp_inverse_of_lane_products (warp_util.cuh) makes all lanes invert the SAME value.
template <class P>
__host__ __device__ __noinline__ fp_t fp_inv(const fp_t a) {
    if (fp_is_zero(a)) return a;
    uint32_t u[8], v[8], r[8], s[8], pl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u[i] = pl[i] = P::p(i);
        v[i] = a.l[i];
        r[i] = 0;
        s[i] = (i == 0);
    }
    uint32_t k = 0;
    // invariants (Kaliski 1995): a~ r == -u 2^k, a~ s == v 2^k (mod p); r, s < 2p < 2^255.
    // The textbook loop does one halving per step (u even: u/2, 2s | v even: v/2, 2r | u > v: (u-v)/2, r+s, 2s |
    // else (v-u)/2, s+r, 2r).  Here u and v are kept ODD: after a subtraction ALL trailing zeros go at once (the same
    // sequence of textbook steps, merged), which halves the number of iterations and drops the per-step parity tests.
    strip_zeros8(v, r, k);  // (r = 0: only v and k change)
    for (;;) {
        uint32_t t[8];
        uint32_t borrow = sub8(t, u, v);
        if (!borrow && (t[0] | t[1] | t[2] | t[3] | t[4] | t[5] | t[6] | t[7]) != 0) {  // u > v
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = t[i];
            add8(r, r, s);
            strip_zeros8(u, s, k);
        } else {
            sub8(v, v, u);
            add8(s, s, r);
            if ((v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) == 0) {  // u == v: the textbook's last step
                shl1_8(r);
                ++k;
                break;
            }
            strip_zeros8(v, r, k);
        }
    }
    if (geq8(r, pl)) sub8(r, r, pl);
    sub8(r, pl, r);  // r = a~^-1 2^k mod p, in [1, p]
    fp_t x, r2, t;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x.l[i] = r[i];
        r2.l[i] = P::r2(i);
    }
    if (geq8(x.l, pl)) sub8(x.l, x.l, pl);
    // x * 2^(512 - k), exponent in [4, 258]: plain product x * y == mont_mul(x, to_mont(y)), to_mont(y) = mont_mul(R^2, y)
    uint32_t j = 512 - k;
    uint32_t j1 = j > 253 ? 253 : j;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.l[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if ((uint32_t)i == (j1 >> 5)) t.l[i] = 1u << (j1 & 31);
    x = fp_mul_ni<P>(x, fp_mul_ni<P>(r2, t));
    if (j > 253) {
        uint32_t j2 = j - 253;  // <= 5
#pragma unroll
        for (int i = 0; i < 8; ++i) t.l[i] = (i == 0) ? (1u << j2) : 0;
        x = fp_mul_ni<P>(x, fp_mul_ni<P>(r2, t));
    }
    return x;
}

// 16-byte vector load/store of a field element (global or shared, 16-byte aligned)
VK_HD fp_t fp_load(const fp_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fp_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ fp_t fp_load_ro(const fp_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fp_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
VK_HD void fp_store(fp_t* p, const fp_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

}  // namespace vk
