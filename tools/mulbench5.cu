// Last candidate for a multiplier that is not a carry chain: 9 x 29-bit limbs,
// Montgomery radix R' = 2^261, every multiply-accumulate a PLAIN mad.wide.u32 into 64-bit column accumulators (18 products of
// < 2^58 each fit), carries moved by shifts once per column — and, unlike tools/mulbench.cu variant 1, with the normalisation
// kept to 4 instructions per column (~95 ALU instructions per product instead of ~170), the modulus as immediates and the
// a x b columns issued before the reduction rows so that a thread offers the multiply pipe up to 17 independent chains.
// Measured like the kernels use a multiplier: a dependent chain per thread, 128-thread CTAs, 1 .. 8 warps per sub-partition,
// (a) the product alone and (b) the product mix of a mixed addition (9 products + 7 additions / subtractions per step).
// Prints G products/s for the 8 x 32-bit multiplier of field.cuh and for the 29-bit candidate, plus a few (a, b, result)
// triples that tools/mulbench5_check.py verifies with Python integers.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mulbench5 tools/mulbench5.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../verkle_kzg_b200/csrc/field.cuh"
using namespace vk;

struct fq29 {
    uint32_t l[9];
};
static const uint32_t M29 = (1u << 29) - 1;

// BN254 Fq modulus in 29-bit limbs and -p^-1 mod 2^29 (checked by mulbench5_check.py)
#define P29_0 0x187cfd47u
#define P29_1 0x010460b6u
#define P29_2 0x1c72a34fu
#define P29_3 0x02d522d0u
#define P29_4 0x1585d978u
#define P29_5 0x02db40c0u
#define P29_6 0x00a6e141u
#define P29_7 0x0e5c2634u
#define P29_8 0x0030644eu
#define P29_INV 0x04866389u

__host__ __device__ inline fq29 to29(const fp_t& a) {
    fq29 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        int bit = 29 * i, w = bit >> 5, s = bit & 31;
        uint64_t v = a.l[w];
        if (w + 1 < 8) v |= (uint64_t)a.l[w + 1] << 32;
        r.l[i] = (uint32_t)(v >> s) & M29;
    }
    return r;
}
__host__ __device__ inline fp_t from29(const fq29& a) {  // limbs normalised, value < 2^256
    fp_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        int bit = 29 * i, w = bit >> 5, s = bit & 31;
        uint64_t v = (uint64_t)a.l[i] << s;
        r.l[w] |= (uint32_t)v;
        if (w + 1 < 8) r.l[w + 1] |= (uint32_t)(v >> 32);
    }
    return r;
}

__device__ __forceinline__ uint64_t mac(uint64_t acc, uint32_t a, uint32_t b) {
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
    return acc;
}

// a b 2^-261 mod p, limbs of the inputs < 2^29 (+ a little slack), output limbs < 2^29, value < 2p for inputs < 2p
__device__ __forceinline__ fq29 mul29(const fq29& a, const fq29& b) {
    const uint32_t P[9] = {P29_0, P29_1, P29_2, P29_3, P29_4, P29_5, P29_6, P29_7, P29_8};
    uint64_t t[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) t[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) t[i + j] = mac(t[i + j], a.l[j], b.l[i]);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint32_t m = ((uint32_t)t[i] * P29_INV) & M29;
#pragma unroll
        for (int j = 0; j < 9; ++j) t[i + j] = mac(t[i + j], m, P[j]);
        t[i + 1] += t[i] >> 29;
    }
    fq29 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        r.l[k] = (uint32_t)t[9 + k] & M29;
        t[10 + k] += t[9 + k] >> 29;
    }
    r.l[8] = (uint32_t)t[17];
    return r;
}
__device__ __noinline__ fq29 mul29_ni(const fq29 a, const fq29 b) { return mul29(a, b); }

// lazy subtraction with headroom (R' = 2^261 leaves 7 bits): a - b + 4p, limb-wise, no carry chain, no comparison.
// 4p in a redundant form whose limbs all exceed 2^29 so that no limb goes negative (limbs of a, b < 2^29 + slack).
__device__ __forceinline__ fq29 sub29(const fq29& a, const fq29& b) {
    // 4p = sum c_i 2^(29 i) with c_i = 4 P_i rewritten: borrow 2^30 from the limb above into each limb but the top one
    const uint32_t P[9] = {P29_0, P29_1, P29_2, P29_3, P29_4, P29_5, P29_6, P29_7, P29_8};
    fq29 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        uint32_t c = 4 * P[i] + (i < 8 ? (1u << 30) : 0u) - (i > 0 ? 2u : 0u);
        r.l[i] = a.l[i] - b.l[i] + c;
    }
    return r;
}
// one carry pass: limbs back below 2^29 (+1 carry bit of slack in the top limb)
__device__ __forceinline__ fq29 norm29(const fq29& a) {
    fq29 r;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t v = a.l[i] + c;
        r.l[i] = v & M29;
        c = v >> 29;
    }
    r.l[8] = a.l[8] + c;
    return r;
}

template <int KIND, bool NI>
__global__ void __launch_bounds__(128) kern(fp_t* x, const fp_t* y, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) {  // 8 x 32: the production multiplier, dependent chain
        fp_t a = x[i], b = y[i];
#pragma unroll 1
        for (int k = 0; k < iters; ++k) a = NI ? fp_mul_lazy_ni<FqParams>(a, b) : fp_mul_lazy<FqParams>(a, b);
        x[i] = a;
    } else if (KIND == 1) {  // 9 x 29 candidate, dependent chain
        fq29 a = to29(x[i]), b = to29(y[i]);
#pragma unroll 1
        for (int k = 0; k < iters; ++k) a = NI ? mul29_ni(a, b) : mul29(a, b);
        x[i] = from29(a);
    } else if (KIND == 2) {  // 8 x 32: the product mix of a mixed addition (9 calls, one of them the fused pair, 7 lazy add / sub)
        fp_t X = x[i], Y = y[i], ZZ = x[i], ZZZ = y[i];
        const fp_t px = y[i], py = x[i];
#pragma unroll 1
        for (int k = 0; k < iters; ++k) {
            fp_t U2 = fp_mul_lazy_ni<FqParams>(px, ZZ), S2 = fp_mul_lazy_ni<FqParams>(py, ZZZ);
            fp_t P = fp_sub_lazy<FqParams>(U2, X), R = fp_sub_lazy<FqParams>(S2, Y);
            fp_t PP = fp_mul_lazy_ni<FqParams>(P, P), PPP = fp_mul_lazy_ni<FqParams>(P, PP), Q = fp_mul_lazy_ni<FqParams>(X, PP);
            fp_t X3 = fp_sub_lazy<FqParams>(fp_sub_lazy<FqParams>(fp_mul_lazy_ni<FqParams>(R, R), PPP), fp_add_lazy<FqParams>(Q, Q));
            Y = fp_mul2_lazy_ni<FqParams>(R, fp_sub_lazy<FqParams>(Q, X3), fp_neg_lazy<FqParams>(Y), PPP);
            X = X3;
            ZZ = fp_mul_lazy_ni<FqParams>(ZZ, PP);
            ZZZ = fp_mul_lazy_ni<FqParams>(ZZZ, PPP);
        }
        x[i] = X;
        x[i].l[0] ^= Y.l[0] ^ ZZ.l[0] ^ ZZZ.l[0];
    } else {  // 9 x 29: the same mix (10 products: no fused pair here), subtractions without carry chains, one carry pass each
        fq29 X = to29(x[i]), Y = to29(y[i]), ZZ = X, ZZZ = Y;
        const fq29 px = Y, py = X;
#pragma unroll 1
        for (int k = 0; k < iters; ++k) {
            fq29 U2 = mul29_ni(px, ZZ), S2 = mul29_ni(py, ZZZ);
            fq29 P = norm29(sub29(U2, X)), R = norm29(sub29(S2, Y));
            fq29 PP = mul29_ni(P, P), PPP = mul29_ni(P, PP), Q = mul29_ni(X, PP);
            fq29 Q2;
#pragma unroll
            for (int j = 0; j < 9; ++j) Q2.l[j] = 2 * Q.l[j];
            fq29 X3 = norm29(sub29(norm29(sub29(mul29_ni(R, R), PPP)), Q2));
            fq29 Y3 = norm29(sub29(mul29_ni(R, norm29(sub29(Q, X3))), mul29_ni(Y, PPP)));
            X = X3;
            Y = Y3;
            ZZ = mul29_ni(ZZ, PP);
            ZZZ = mul29_ni(ZZZ, PPP);
        }
        x[i] = from29(X);
        x[i].l[0] ^= Y.l[0] ^ ZZ.l[0] ^ ZZZ.l[0];
    }
}

template <int KIND, bool NI>
static void run(const char* name, int muls_per_iter, fp_t* dx, fp_t* dy, const std::vector<fp_t>& hx) {
    cudaFuncSetAttribute(kern<KIND, NI>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern<KIND, NI>);
    for (int wps : {1, 2, 4, 8}) {
        int blocks = 148 * wps;  // 128-thread CTAs: one warp per sub-partition each
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<KIND, NI>, 128, 0);
        if (occ < wps) {
            printf("%-44s warps/SMSP %d: occupancy allows %d CTAs/SM (regs %d) - skipped\n", name, wps, occ, fa.numRegs);
            continue;
        }
        int n = blocks * 128;
        cudaMemcpy(dx, hx.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
        int iters = KIND < 2 ? 4000 : 400;
        kern<KIND, NI><<<blocks, 128>>>(dx, dy, 50);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaMemcpy(dx, hx.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        kern<KIND, NI><<<blocks, 128>>>(dx, dy, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double g = (double)n * iters * muls_per_iter / ms / 1e6;
        printf("%-44s warps/SMSP %d regs %3d: %8.3f ms  %7.2f G products/s  (%s)\n", name, wps, fa.numRegs, ms, g,
               cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    const int nmax = 148 * 8 * 128;
    std::vector<fp_t> hx(nmax), hy(nmax);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    auto rnd = [&] {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return (uint32_t)(s >> 16);
    };
    for (int i = 0; i < nmax; ++i) {
        for (int k = 0; k < 8; ++k) { hx[i].l[k] = rnd(); hy[i].l[k] = rnd(); }
        hx[i].l[7] &= 0x1fffffffu;  // < 2^253 < p
        hy[i].l[7] &= 0x1fffffffu;
    }
    fp_t *dx, *dy;
    cudaMalloc(&dx, nmax * sizeof(fp_t));
    cudaMalloc(&dy, nmax * sizeof(fp_t));
    cudaMemcpy(dy, hy.data(), nmax * sizeof(fp_t), cudaMemcpyHostToDevice);
    // correctness triples: one product each
    cudaMemcpy(dx, hx.data(), 128 * sizeof(fp_t), cudaMemcpyHostToDevice);
    kern<1, false><<<1, 128>>>(dx, dy, 1);
    std::vector<fp_t> out(4);
    cudaMemcpy(out.data(), dx, 4 * sizeof(fp_t), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 4; ++i) {
        printf("triple");
        for (const fp_t* v : {&hx[i], &hy[i], &out[i]}) {
            printf(" ");
            for (int k = 7; k >= 0; --k) printf("%08x", v->l[k]);
        }
        printf("\n");
    }
    run<0, false>("8x32 carry chains, inlined, product chain", 1, dx, dy, hx);
    run<0, true>("8x32 carry chains, called, product chain", 1, dx, dy, hx);
    run<1, false>("9x29 plain mad.wide, inlined, product chain", 1, dx, dy, hx);
    run<1, true>("9x29 plain mad.wide, called, product chain", 1, dx, dy, hx);
    run<2, true>("8x32 mixed-addition mix (9 calls incl. fused)", 10, dx, dy, hx);
    run<3, true>("9x29 mixed-addition mix (10 calls)", 10, dx, dy, hx);
    return 0;
}
