// Karatsuba form of the lazy Montgomery product: a measured NEGATIVE result, kept out of field.cuh so that the header the
// kernels compile holds only what they call.  Included by api.cu (the probe, vkzg_probe_fq_sqr_dev modes 2 / 4) and by the host
// harness (tests/host/host_check.cu); tests: tests/test_host_arith.py::test_karatsuba_product, tests/test_gpu_field.py.
#pragma once
#include "field.cuh"

namespace vk {

// ---------------------------------------------------------------------------------------------
// Karatsuba form of the lazy product — a measured NEGATIVE result, kept with its tests and the probe (vkzg_probe_fq_sqr_dev
// modes 2 / 4) but not called by any kernel: 113 instead of 129 multiplier-pipe instructions per product, yet 66.80 against
// 66.85 G products/s on B200 (profiles/r02_karatsuba_probe.txt): the ~120 extra carry-chain instructions of the glue are not
// free on this SM (each costs about an eighth of a carry-chained IMAD.WIDE), where the dedicated square's 44 were.
//   a = a1 2^128 + a0, b = b1 2^128 + b0:   a b = z0 + (z0 + z2 - (a1 - a0)(b1 - b0)) 2^128 + z2 2^256
// three 4 x 4-limb products (48 multiply-accumulates instead of 64) glued by carry chains on the ALU pipe, then the same
// eight reduction rows as the interleaved form, run over the low half T_lo with no product rows in between; the high half
// is added at the end:  (T_lo + M p) / R + T_hi < p + 1 + 4 p^2 / R < 2p.
// ---------------------------------------------------------------------------------------------
// r[0..7] = x[0..3] * y[0..3]: row 0 as plain products, rows 1..3 as two 2-pair carry chains each (even / odd columns)
VK_HD void mul4x4(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint32_t E[8], O[8];  // E: pairs at columns (0,1)(2,3)(4,5)(6,7);  O[k] = column k: pairs (1,2)(3,4)(5,6), O[7] a carry
    {
        uint64_t t0 = (uint64_t)x[0] * y[0], t1 = (uint64_t)x[2] * y[0], t2 = (uint64_t)x[1] * y[0], t3 = (uint64_t)x[3] * y[0];
        E[0] = (uint32_t)t0; E[1] = (uint32_t)(t0 >> 32); E[2] = (uint32_t)t1; E[3] = (uint32_t)(t1 >> 32);
        O[1] = (uint32_t)t2; O[2] = (uint32_t)(t2 >> 32); O[3] = (uint32_t)t3; O[4] = (uint32_t)(t3 >> 32);
    }
#ifdef __CUDA_ARCH__
    // row 1 (y1): O (1,2) += x0 y1, (3,4) += x2 y1, carry -> O5 ;  E (2,3) += x1 y1, (4,5) = x3 y1 + carry
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %13, %3;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "mad.lo.cc.u32 %5, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %6, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "=r"(O[5]), "+r"(E[2]), "+r"(E[3]), "=r"(E[4]), "=r"(E[5])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[1]));
    // row 2 (y2): E (2,3) += x0 y2, (4,5) += x2 y2, carry -> E6 ;  O (3,4) += x1 y2, (5,6) = x3 y2 + O5 + carry
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %13, %3;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "mad.lo.cc.u32 %5, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %6, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, %7;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "=r"(E[6]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "=r"(O[6])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[2]));
    // row 3 (y3): O (3,4) += x0 y3, (5,6) += x2 y3, carry -> O7 ;  E (4,5) += x1 y3, (6,7) = x3 y3 + E6 + carry
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %13, %3;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "mad.lo.cc.u32 %5, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %6, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, %7;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "=r"(O[7]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "=r"(E[7])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[3]));
    // r = E + (O << 32)
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]),
          "r"(O[5]), "r"(O[6]), "r"(O[7]));
    r[0] = E[0];
#else
    uint32_t c = 0, z = 0;
    // row 1
    host_mad_pair(O[1], O[2], x[0], y[1], O[1], O[2], c);
    host_mad_pair(O[3], O[4], x[2], y[1], O[3], O[4], c);
    O[5] = c; c = 0;
    host_mad_pair(E[2], E[3], x[1], y[1], E[2], E[3], c);
    host_mad_pair(E[4], E[5], x[3], y[1], z, z, c);
    // row 2
    c = 0;
    host_mad_pair(E[2], E[3], x[0], y[2], E[2], E[3], c);
    host_mad_pair(E[4], E[5], x[2], y[2], E[4], E[5], c);
    E[6] = c; c = 0;
    host_mad_pair(O[3], O[4], x[1], y[2], O[3], O[4], c);
    host_mad_pair(O[5], O[6], x[3], y[2], O[5], z, c);
    // row 3
    c = 0;
    host_mad_pair(O[3], O[4], x[0], y[3], O[3], O[4], c);
    host_mad_pair(O[5], O[6], x[2], y[3], O[5], O[6], c);
    O[7] = c; c = 0;
    host_mad_pair(E[4], E[5], x[1], y[3], E[4], E[5], c);
    host_mad_pair(E[6], E[7], x[3], y[3], E[6], z, c);
    uint64_t cy = 0;
    r[0] = E[0];
    for (int k = 1; k < 8; ++k) {
        cy += (uint64_t)E[k] + O[k];
        r[k] = (uint32_t)cy;
        cy >>= 32;
    }
#endif
}

// d = |x - y| over 4 limbs, returns 0xffffffff if x < y else 0
VK_HD uint32_t absdiff4(uint32_t* d, const uint32_t* x, const uint32_t* y) {
    uint32_t neg;
#ifdef __CUDA_ARCH__
    uint32_t t0, t1, t2, t3;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(neg)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
    // (t xor neg) - neg
    t0 ^= neg; t1 ^= neg; t2 ^= neg; t3 ^= neg;
    asm("sub.cc.u32 %0, %4, %8;\n\t"
        "subc.cc.u32 %1, %5, %8;\n\t"
        "subc.cc.u32 %2, %6, %8;\n\t"
        "subc.u32 %3, %7, %8;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(neg));
#else
    uint32_t t[4], bw = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = (uint64_t)x[i] - y[i] - bw;
        t[i] = (uint32_t)v;
        bw = (uint32_t)(v >> 32) & 1;
    }
    neg = bw ? 0xffffffffu : 0u;
    bw = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = (uint64_t)(t[i] ^ neg) - neg - bw;
        d[i] = (uint32_t)v;
        bw = (uint32_t)(v >> 32) & 1;
    }
#endif
    return neg;
}

// T[0..15] = a * b by one level of Karatsuba
VK_HD void mul8x8_karatsuba(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[8], z2[8], m[8], da[4], db[4];
    mul4x4(z0, a, b);
    mul4x4(z2, a + 4, b + 4);
    const uint32_t sa = absdiff4(da, a + 4, a), sb = absdiff4(db, b + 4, b);
    mul4x4(m, da, db);
    // z1 = z0 + z2 - sigma m, sigma = +1 when the two differences have the same sign:  S + (m xor k) + (k & 1) - (k ? 2^256 : 0)
    const uint32_t k = ~(sa ^ sb);  // all ones: subtract m
    uint32_t S[9], z1[9];
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
        : "=r"(S[0]), "=r"(S[1]), "=r"(S[2]), "=r"(S[3]), "=r"(S[4]), "=r"(S[5]), "=r"(S[6]), "=r"(S[7]), "=r"(S[8])
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]),
          "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
    uint32_t mx[8], scratch;
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = m[i] ^ k;
    // carry-in = k & 1: (k + k) sets the carry flag exactly when k is all ones
    asm("add.cc.u32 %9, %26, %26;\n\t"
        "addc.cc.u32 %0, %10, %18;\n\t"
        "addc.cc.u32 %1, %11, %19;\n\t"
        "addc.cc.u32 %2, %12, %20;\n\t"
        "addc.cc.u32 %3, %13, %21;\n\t"
        "addc.cc.u32 %4, %14, %22;\n\t"
        "addc.cc.u32 %5, %15, %23;\n\t"
        "addc.cc.u32 %6, %16, %24;\n\t"
        "addc.cc.u32 %7, %17, %25;\n\t"
        "addc.u32 %8, %27, %26;"
        : "=r"(z1[0]), "=r"(z1[1]), "=r"(z1[2]), "=r"(z1[3]), "=r"(z1[4]), "=r"(z1[5]), "=r"(z1[6]), "=r"(z1[7]), "=r"(z1[8]), "=&r"(scratch)
        : "r"(S[0]), "r"(S[1]), "r"(S[2]), "r"(S[3]), "r"(S[4]), "r"(S[5]), "r"(S[6]), "r"(S[7]), "r"(mx[0]), "r"(mx[1]), "r"(mx[2]), "r"(mx[3]),
          "r"(mx[4]), "r"(mx[5]), "r"(mx[6]), "r"(mx[7]), "r"(k), "r"(S[8]));
    // T = z0 + z1 2^128 + z2 2^256 (two blocks: the operand count of one asm statement is limited)
    uint32_t cmid;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(cmid)
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z1[0]), "r"(z1[1]), "r"(z1[2]), "r"(z1[3]));
    asm("add.cc.u32 %8, %22, 0xffffffff;\n\t"   // restores the carry flag from cmid
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, 0;\n\t"
        "addc.cc.u32 %6, %15, 0;\n\t"
        "addc.u32 %7, %16, 0;"
        : "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15]), "=&r"(scratch)
        : "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(z1[4]), "r"(z1[5]), "r"(z1[6]),
          "r"(z1[7]), "r"(z1[8]), "r"(cmid));
#else
    uint64_t cy = 0;
    for (int i = 0; i < 8; ++i) {
        cy += (uint64_t)z0[i] + z2[i];
        S[i] = (uint32_t)cy;
        cy >>= 32;
    }
    S[8] = (uint32_t)cy;
    cy = k & 1;
    for (int i = 0; i < 8; ++i) {
        cy += (uint64_t)S[i] + (m[i] ^ k);
        z1[i] = (uint32_t)cy;
        cy >>= 32;
    }
    z1[8] = S[8] + (uint32_t)cy + k;
    cy = 0;
    for (int i = 4; i < 16; ++i) {
        cy += (uint64_t)(i < 8 ? z0[i] : z2[i - 8]) + (i - 4 < 9 ? z1[i - 4] : 0u);
        T[i] = (uint32_t)cy;
        cy >>= 32;
    }
#endif
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = z0[i];
}

// one reduction row AFTER a shift of the frame, with no product row in between:
//   x0 += e1 ;  m = x0 * (-p^-1) ;  y[k,k+1] = e[k+2,k+3] + p_odd m (carry of the first addition flows in) ; y[6,7] = p7 m + carry ;
//   x[..] += p_even m, carry into y7
template <class P>
VK_HD void shift_reduce_row(uint32_t* x, uint32_t* y, const uint32_t* e) {
#ifdef __CUDA_ARCH__
    uint32_t m;
    asm("add.cc.u32 %0, %0, %10;\n\t"
        "mul.lo.u32 %9, %0, %17;\n\t"
        "madc.lo.cc.u32 %1, %18, %9, %11;\n\t"
        "madc.hi.cc.u32 %2, %18, %9, %12;\n\t"
        "madc.lo.cc.u32 %3, %19, %9, %13;\n\t"
        "madc.hi.cc.u32 %4, %19, %9, %14;\n\t"
        "madc.lo.cc.u32 %5, %20, %9, %15;\n\t"
        "madc.hi.cc.u32 %6, %20, %9, %16;\n\t"
        "madc.lo.cc.u32 %7, %21, %9, 0;\n\t"
        "madc.hi.u32 %8, %21, %9, 0;"
        : "+r"(x[0]), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7]), "=&r"(m)
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(P::INV), "r"(P::p(1)), "r"(P::p(3)), "r"(P::p(5)),
          "r"(P::p(7)));
    mad_row4(x, y[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
#else
    uint64_t s = (uint64_t)x[0] + e[1];
    x[0] = (uint32_t)s;
    uint32_t c = (uint32_t)(s >> 32);
    const uint32_t m = x[0] * P::INV;
    host_mad_pair(y[0], y[1], P::p(1), m, e[2], e[3], c);
    host_mad_pair(y[2], y[3], P::p(3), m, e[4], e[5], c);
    host_mad_pair(y[4], y[5], P::p(5), m, e[6], e[7], c);
    host_mad_pair(y[6], y[7], P::p(7), m, 0, 0, c);
    mad_row4(x, y[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
#endif
}

// a b R^-1 mod p in [0, 2p) for a, b in [0, 2p]
template <class P>
VK_HD fp_t fp_mul_lazy_kara(const fp_t& a, const fp_t& b) {
    uint32_t T[16], u[8], v[8], y[8];
    mul8x8_karatsuba(T, a.l, b.l);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        u[k] = T[k];
        v[k] = 0;
    }
    reduce_step<P>(u, v);
#define VK_RED_ROW(X, E)               \
    shift_reduce_row<P>(X, y, E);      \
    _Pragma("unroll") for (int k = 0; k < 8; ++k) E[k] = y[k];
    VK_RED_ROW(v, u)
    VK_RED_ROW(u, v)
    VK_RED_ROW(v, u)
    VK_RED_ROW(u, v)
    VK_RED_ROW(v, u)
    VK_RED_ROW(u, v)
    VK_RED_ROW(v, u)
#undef VK_RED_ROW
    uint32_t vs[8], lo[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) vs[k] = v[k + 1];
    vs[7] = 0;
    add8(lo, u, vs);
    fp_t r;
    add8(r.l, lo, T + 8);
    return r;
}
template <class P>
__host__ __device__ __noinline__ fp_t fp_mul_lazy_kara_ni(const fp_t a, const fp_t b) {
    return fp_mul_lazy_kara<P>(a, b);
}

}  // namespace vk
