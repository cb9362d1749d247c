// Instruction-level microbenchmarks for the multiplier design (sm_100a).  Every kernel is a ring of dependent
// operations that ptxas cannot reassociate (the multiplicands are other chains' results), so the SASS of the loop
// body is exactly the instruction mix named here — check with
//   cuobjdump -sass tools/ubench2 | less
// Timing is in SM cycles (clock64 inside the kernel), so DVFS does not matter.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define ROUNDS 8
enum { WIDE_RZ, WIDE_ACC, WIDE_ACC_IMM, WIDE_X, WIDE_RZ_IADD3, IMAD32, DFMA, WIDE_RZ_DFMA, WIDE_RZ_ALU1, WIDE_RZ_ALU2, WIDE_ACC_ALU1,
       DFMA_ALU1, ALU_ONLY, IMADHI, WIDE_X_ALU1, WIDE_CCOUT, DFMA_IADD64, WIDE_ACC_ALU2, WIDE_X_IMM, NKINDS };
static const char* names[] = {"IMAD.WIDE rz (product only)", "IMAD.WIDE acc (64-bit addend, 4 src regs)", "IMAD.WIDE acc, immediate multiplicand",
                              "IMAD.WIDE .cc/.X chain of 4 (as fp_mul rows)", "2 IMAD.WIDE rz + IADD3/IADD3.X 3-input 64-bit add", "IMAD 32-bit acc",
                              "DFMA", "IMAD.WIDE rz + DFMA 1:1", "IMAD.WIDE rz + 1 LOP3", "IMAD.WIDE rz + 2 LOP3", "IMAD.WIDE acc + 1 LOP3",
                              "DFMA + 1 LOP3", "LOP3 only", "IMAD.HI acc", "X chain of 4 + 4 LOP3", "IMAD.WIDE acc with carry-out (no carry-in) + addc capture",
                              "2 DFMA + DADD + 2x(IADD3+IADD3.X) (Emmart limb product)", "IMAD.WIDE acc + 2 LOP3", "X chain of 4, immediate multiplicands"};

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t iters, const uint32_t* seed, uint64_t* sink, long long* cyc) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s[16], b[8], l[8];
    double d[8], dc[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = seed[j] + t * 2654435761u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        b[j] = seed[16 + j] ^ t;
        l[j] = seed[24 + j] + t;
        d[j] = 1.0 + (seed[j] & 0xffff) * 1e-9;
        dc[j] = 1.0 - (seed[8 + j] & 0xffff) * 1e-12;
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int a0 = (2 * j + 2) & 15, a1 = (2 * j + 3) & 15;
                if (KIND == WIDE_RZ || KIND == WIDE_RZ_DFMA || KIND == WIDE_RZ_ALU1 || KIND == WIDE_RZ_ALU2) {
                    uint64_t p = (uint64_t)s[a0] * s[a1];
                    s[2 * j] = (uint32_t)p;
                    s[2 * j + 1] = (uint32_t)(p >> 32);
                }
                if (KIND == WIDE_ACC || KIND == WIDE_ACC_ALU1 || KIND == WIDE_ACC_ALU2) {
                    uint64_t acc = ((uint64_t)s[2 * j + 1] << 32) | s[2 * j];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(s[a0]), "r"(b[j]));
                    s[2 * j] = (uint32_t)acc;
                    s[2 * j + 1] = (uint32_t)(acc >> 32);
                }
                if (KIND == WIDE_ACC_IMM) {
                    uint64_t acc = ((uint64_t)s[2 * j + 1] << 32) | s[2 * j];
                    asm volatile("mad.wide.u32 %0, %1, 0x3c208c16, %0;" : "+l"(acc) : "r"(s[a0]));
                    s[2 * j] = (uint32_t)acc;
                    s[2 * j + 1] = (uint32_t)(acc >> 32);
                }
                if (KIND == IMAD32) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s[2 * j]) : "r"(s[a0]), "r"(b[j]));
                if (KIND == IMADHI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(s[2 * j]) : "r"(s[a0]), "r"(b[j]));
                if (KIND == DFMA || KIND == WIDE_RZ_DFMA || KIND == DFMA_ALU1)
                    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(d[(j + 1) & 7]), "d"(dc[j]));
                if (KIND == WIDE_RZ_ALU1 || KIND == WIDE_RZ_ALU2 || KIND == WIDE_ACC_ALU1 || KIND == WIDE_ACC_ALU2 || KIND == DFMA_ALU1 || KIND == ALU_ONLY)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[j]) : "r"(l[(j + 1) & 7]), "r"(l[(j + 3) & 7]));
                if (KIND == WIDE_RZ_ALU2 || KIND == WIDE_ACC_ALU2 || KIND == ALU_ONLY)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x69;" : "+r"(l[(j + 4) & 7]) : "r"(l[(j + 5) & 7]), "r"(l[(j + 6) & 7]));
                if (KIND == WIDE_CCOUT) {
                    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                                 : "+r"(s[2 * j]), "+r"(s[2 * j + 1]), "+r"(l[j])
                                 : "r"(s[a0]), "r"(b[j]));
                }
                if (KIND == DFMA_IADD64) {
                    // hi = fma_rz(a, b, c1); lo = fma_rz(a, b, c2 - hi); two 64-bit integer accumulations of the bit patterns
                    double hi, lo, sub;
                    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(d[(j + 1) & 7]), "d"(dc[j]), "d"(dc[(j + 1) & 7]));
                    asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(sub) : "d"(dc[(j + 2) & 7]), "d"(hi));
                    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(d[(j + 1) & 7]), "d"(dc[j]), "d"(sub));
                    uint64_t acc0 = ((uint64_t)s[2 * j + 1] << 32) | s[2 * j];
                    acc0 += (uint64_t)__double_as_longlong(hi);
                    s[2 * j] = (uint32_t)acc0;
                    s[2 * j + 1] = (uint32_t)(acc0 >> 32);
                    uint64_t acc1 = ((uint64_t)l[j] << 32) | b[j];
                    acc1 += (uint64_t)__double_as_longlong(lo);
                    b[j] = (uint32_t)acc1;
                    l[j] = (uint32_t)(acc1 >> 32);
                    d[j] = __longlong_as_double((long long)((acc0 & 0x000fffffffffffffull) | 0x3ff0000000000000ull));
                }
            }
            if (KIND == WIDE_RZ_IADD3) {
                // 8 products (RZ) -> 4 accumulators, each += two products with one IADD3 / IADD3.X pair
                uint64_t p[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) p[j] = (uint64_t)s[(2 * j + 2) & 15] * b[j];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t acc = ((uint64_t)s[4 * j + 1] << 32) | s[4 * j];
                    acc += p[2 * j] + p[2 * j + 1];
                    s[4 * j] = (uint32_t)acc;
                    s[4 * j + 1] = (uint32_t)(acc >> 32);
                    // keep the ring moving: odd accumulators are rewritten from the even ones
                    s[4 * j + 2] ^= s[4 * j];
                }
            }
            if (KIND == WIDE_X || KIND == WIDE_X_ALU1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int o = 8 * h;
                    asm volatile(
                        "mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                        "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                        "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                        "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
                        : "+r"(s[o + 0]), "+r"(s[o + 1]), "+r"(s[o + 2]), "+r"(s[o + 3]), "+r"(s[o + 4]), "+r"(s[o + 5]), "+r"(s[o + 6]), "+r"(s[o + 7])
                        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(s[(o + 8 + r) & 15]));
                    if (KIND == WIDE_X_ALU1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[j + 4 * h]) : "r"(l[(j + 1 + 4 * h) & 7]), "r"(l[(j + 3 + 4 * h) & 7]));
                    }
                }
            }
            if (KIND == WIDE_X_IMM) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int o = 8 * h;
                    asm volatile(
                        "mad.lo.cc.u32 %0, %8, 0x3c208c16, %0;\n\tmadc.hi.cc.u32 %1, %8, 0x3c208c16, %1;\n\t"
                        "madc.lo.cc.u32 %2, %8, 0x97816a91, %2;\n\tmadc.hi.cc.u32 %3, %8, 0x97816a91, %3;\n\t"
                        "madc.lo.cc.u32 %4, %8, 0xb85045b6, %4;\n\tmadc.hi.cc.u32 %5, %8, 0xb85045b6, %5;\n\t"
                        "madc.lo.cc.u32 %6, %8, 0x30644e72, %6;\n\tmadc.hi.u32 %7, %8, 0x30644e72, %7;"
                        : "+r"(s[o + 0]), "+r"(s[o + 1]), "+r"(s[o + 2]), "+r"(s[o + 3]), "+r"(s[o + 4]), "+r"(s[o + 5]), "+r"(s[o + 6]), "+r"(s[o + 7])
                        : "r"(s[(o + 8 + r) & 15]));
                }
            }
        }
    }
    long long t1 = clock64();
    uint64_t x = 0;
    double ds = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x ^= s[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x ^= l[j] ^ b[j]; ds += d[j]; }
    if (x == 0x123456789abcdefull || ds == 1234.5678) sink[0] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static int instr_per_round(int kind) {  // "primary" operations per round (what the cycles are divided by)
    switch (kind) {
        case WIDE_X: case WIDE_X_ALU1: case WIDE_X_IMM: return 8;
        default: return 8;
    }
}

template <int KIND>
void run(uint32_t* d_seed) {
    uint64_t* sink;
    long long* d_cyc;
    cudaMalloc(&sink, 8);
    for (int wps : {4, 8, 16}) {                 // warps per SM sub-partition
        int blocks = 148 * wps / 2;              // 256 threads = 8 warps = 2 per SMSP
        cudaMalloc(&d_cyc, blocks * sizeof(long long));
        uint32_t iters = 2000;
        k<KIND><<<blocks, 256>>>(50, d_seed, sink, d_cyc);
        k<KIND><<<blocks, 256>>>(iters, d_seed, sink, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(blocks);
        cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= blocks;
        double per = avg / ((double)iters * ROUNDS * instr_per_round(KIND) * wps);
        printf("%-66s %2d warps/SMSP: %6.2f cycles per primary op per SMSP  (%s)\n", names[KIND], wps, per, cudaGetErrorString(e));
        cudaFree(d_cyc);
    }
    cudaFree(sink);
}

int main() {
    uint32_t h_seed[32];
    for (int i = 0; i < 32; ++i) h_seed[i] = 0x9e3779b9u * (i + 1) + 12345;
    uint32_t* d_seed;
    cudaMalloc(&d_seed, sizeof(h_seed));
    cudaMemcpy(d_seed, h_seed, sizeof(h_seed), cudaMemcpyHostToDevice);
    run<WIDE_RZ>(d_seed);
    run<WIDE_ACC>(d_seed);
    run<WIDE_ACC_IMM>(d_seed);
    run<WIDE_X>(d_seed);
    run<WIDE_X_IMM>(d_seed);
    run<WIDE_CCOUT>(d_seed);
    run<WIDE_RZ_IADD3>(d_seed);
    run<IMAD32>(d_seed);
    run<IMADHI>(d_seed);
    run<DFMA>(d_seed);
    run<ALU_ONLY>(d_seed);
    run<WIDE_RZ_DFMA>(d_seed);
    run<WIDE_RZ_ALU1>(d_seed);
    run<WIDE_RZ_ALU2>(d_seed);
    run<WIDE_ACC_ALU1>(d_seed);
    run<WIDE_ACC_ALU2>(d_seed);
    run<WIDE_X_ALU1>(d_seed);
    run<DFMA_ALU1>(d_seed);
    run<DFMA_IADD64>(d_seed);
    return 0;
}
