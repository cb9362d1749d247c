// Pipe-rate microbenchmark on sm_100a: IMAD.WIDE / IMAD / DFMA with shared vs distinct source registers.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pipebench tools/pipebench.cu
#include <cstdio>
#include <cstdint>
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t iters, uint64_t* sink) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], b[8];
    uint64_t acc[8];
    double da[8], db[8], dacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = t * 2654435761u + j * 977 + 1;
        b[j] = (t ^ 0x9e3779b9u) + j * 131;
        acc[j] = j + t;
        da[j] = 1.0 + j * 1e-9 + t * 1e-12;
        db[j] = 1.0 - j * 1e-9;
        dacc[j] = j;
    }
#pragma unroll 1
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[0]), "r"(b[0]));              // shared sources
                if (KIND == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[j]), "r"(b[(j + r) & 7]));    // distinct sources
                if (KIND == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[j]), "r"(b[r]));              // one shared (row of a Montgomery product)
                if (KIND == 3) { uint32_t lo = (uint32_t)acc[j]; asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a[j]), "r"(b[(j + r) & 7])); acc[j] = (acc[j] & 0xffffffff00000000ull) | lo; }
                if (KIND == 4) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[j]) : "d"(da[0]), "d"(db[0]));
                if (KIND == 5) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(dacc[j]) : "d"(da[j]), "d"(db[(j + r) & 7]));
                if (KIND == 6) { uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);                                  // 32-bit accumulate halves: IMAD.LO + IMAD.HI
                    asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a[j]), "r"(b[(j + r) & 7])); acc[j] = ((uint64_t)hi << 32) | lo; }
            }
        }
    }
    uint64_t s = 0;
    double ds = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s ^= acc[j]; ds += dacc[j]; }
    if (s == 0x123456789abcdefull || ds == 1234.5678) sink[0] = s;
}
template <int KIND>
void run(const char* name, int opsPerInner) {
    uint64_t* sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int wps : {8, 16, 32}) {
        int blocks = 148 * (wps / 8);
        uint32_t iters = 4000;
        k<KIND><<<blocks, 256>>>(100, sink);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            k<KIND><<<blocks, 256>>>(iters, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        double ops = (double)blocks * 256 * iters * 64 * opsPerInner;
        printf("%-58s %2d warps/SM: %7.2f T instr/s  (%.2f instr/clk/SM @1.965GHz)\n", name, wps, ops / best / 1e9, ops / best / 1e9 * 1e12 / 148 / 1.965e9);
    }
}
int main() {
    run<0>("mad.wide.u32 shared a,b", 1);
    run<1>("mad.wide.u32 distinct a,b", 1);
    run<2>("mad.wide.u32 distinct a, shared b", 1);
    run<3>("mad.lo.u32 distinct", 1);
    run<4>("fma.f64 shared", 1);
    run<5>("fma.f64 distinct", 1);
    run<6>("mad.lo + mad.hi distinct (2 instr)", 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
