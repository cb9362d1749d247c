// Multiplier variants on the 8 x 32-bit even/odd-column layout of field.cuh: which of the four carry chains of a row
// (a-even, a-odd, m-even, m-odd) are computed as IMAD.WIDE accumulate chains (mad.lo.cc/madc.hi.cc) and which as
// product-only IMAD.WIDE + an IADD3 carry chain on the ALU pipe.  MASK bit k set = chain type k split, ROWMASK selects rows.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mulbench2 tools/mulbench2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../verkle_kzg_b200/csrc/field.cuh"
using namespace vk;

// acc[0..7] += {a0,a2,a4,a6} * b (64-bit products at column pairs), top += carry : product-only + add chain
__device__ __forceinline__ void s_mad_row4(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    uint64_t p0 = (uint64_t)a0 * b, p1 = (uint64_t)a2 * b, p2 = (uint64_t)a4 * b, p3 = (uint64_t)a6 * b;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"((uint32_t)p0), "r"((uint32_t)(p0 >> 32)), "r"((uint32_t)p1), "r"((uint32_t)(p1 >> 32)), "r"((uint32_t)p2), "r"((uint32_t)(p2 >> 32)),
          "r"((uint32_t)p3), "r"((uint32_t)(p3 >> 32)));
}
__device__ __forceinline__ void s_mad_row4_nc(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    uint64_t p0 = (uint64_t)a0 * b, p1 = (uint64_t)a2 * b, p2 = (uint64_t)a4 * b, p3 = (uint64_t)a6 * b;
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"((uint32_t)p0), "r"((uint32_t)(p0 >> 32)), "r"((uint32_t)p1), "r"((uint32_t)(p1 >> 32)), "r"((uint32_t)p2), "r"((uint32_t)(p2 >> 32)),
          "r"((uint32_t)p3), "r"((uint32_t)(p3 >> 32)));
}
// x0 += e1 ; y[0..7] = e[2..7],0,0 + {a1,a3,a5,a7} * b + carry
__device__ __forceinline__ void s_shift_mad_row4(uint32_t& x0, uint32_t* y, const uint32_t* e, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    uint64_t p0 = (uint64_t)a1 * b, p1 = (uint64_t)a3 * b, p2 = (uint64_t)a5 * b, p3 = (uint64_t)a7 * b;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %10, %16;\n\t"
        "addc.cc.u32 %2, %11, %17;\n\t"
        "addc.cc.u32 %3, %12, %18;\n\t"
        "addc.cc.u32 %4, %13, %19;\n\t"
        "addc.cc.u32 %5, %14, %20;\n\t"
        "addc.cc.u32 %6, %15, %21;\n\t"
        "addc.cc.u32 %7, %22, 0;\n\t"
        "addc.u32 %8, %23, 0;"
        : "+r"(x0), "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"((uint32_t)p0), "r"((uint32_t)(p0 >> 32)), "r"((uint32_t)p1),
          "r"((uint32_t)(p1 >> 32)), "r"((uint32_t)p2), "r"((uint32_t)(p2 >> 32)), "r"((uint32_t)p3), "r"((uint32_t)(p3 >> 32)));
}

template <class P, int MASK, int ROWMASK>
__device__ __forceinline__ void row(uint32_t* x, uint32_t* y, const uint32_t* e, const fp_t& a, uint32_t b, int i) {
    const bool sel = (ROWMASK >> i) & 1;
    // x: column-0 array (its x[0] absorbs e[1]); y: new column-1 array built from e >> 64 + odd products
    if (sel && (MASK & 2)) s_shift_mad_row4(x[0], y, e, a.l[1], a.l[3], a.l[5], a.l[7], b);
    else shift_mad_row4(x[0], y, e, a.l[1], a.l[3], a.l[5], a.l[7], b);
    if (sel && (MASK & 1)) s_mad_row4(x, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b);
    else mad_row4(x, y[7], a.l[0], a.l[2], a.l[4], a.l[6], b);
    uint32_t m = x[0] * P::INV;
    if (sel && (MASK & 8)) s_mad_row4_nc(y, P::p(1), P::p(3), P::p(5), P::p(7), m);
    else mad_row4_nc(y, P::p(1), P::p(3), P::p(5), P::p(7), m);
    if (sel && (MASK & 4)) s_mad_row4(x, y[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
    else mad_row4(x, y[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
}

template <class P, int MASK, int ROWMASK>
__device__ __forceinline__ fp_t mulv(const fp_t& a, const fp_t& b) {
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
    {
        uint32_t m = u[0] * P::INV;
        if ((ROWMASK & 1) && (MASK & 8)) s_mad_row4_nc(v, P::p(1), P::p(3), P::p(5), P::p(7), m);
        else mad_row4_nc(v, P::p(1), P::p(3), P::p(5), P::p(7), m);
        if ((ROWMASK & 1) && (MASK & 4)) s_mad_row4(u, v[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
        else mad_row4(u, v[7], P::p(0), P::p(2), P::p(4), P::p(6), m);
    }
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        {
            uint32_t y[8];
            row<P, MASK, ROWMASK>(v, y, u, a, b.l[i], i);
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = y[k];
        }
        if (i + 1 < 8) {
            uint32_t y[8];
            row<P, MASK, ROWMASK>(u, y, v, a, b.l[i + 1], i + 1);
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

template <int MASK, int ROWMASK>
__global__ void __launch_bounds__(256) kk(fp_t* x, const fp_t* y, int n, int iters, long long* cyc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    fp_t a = x[i], b = y[i];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int k = 0; k < iters; ++k) a = mulv<FqParams, MASK, ROWMASK>(a, b);
    long long t1 = clock64();
    x[i] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static std::vector<fp_t> g_ref;
template <int MASK, int ROWMASK>
void run(fp_t* dx, fp_t* dy, const std::vector<fp_t>& hx, long long* d_cyc) {
    for (int tpsm : {1024, 2048}) {
        int n = 148 * tpsm, blocks = n / 256;
        cudaMemcpy(dx, hx.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
        int iters = 1000;
        kk<MASK, ROWMASK><<<blocks, 256>>>(dx, dy, n, iters, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(blocks);
        cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= blocks;
        std::vector<fp_t> out(64);
        cudaMemcpy(out.data(), dx, 64 * sizeof(fp_t), cudaMemcpyDeviceToHost);
        bool same = true;
        if (g_ref.empty()) g_ref = out;
        else for (int i = 0; i < 64; ++i) for (int k = 0; k < 8; ++k) same &= out[i].l[k] == g_ref[i].l[k];
        // cycles per warp-multiplication per SM sub-partition
        double per = avg / ((double)iters * (tpsm / 32 / 4));
        printf("mask %2d rows 0x%02x  threads/SM %4d : %7.1f cycles per warp-mul per SMSP  -> %6.1f G mul/s @1.965GHz  same=%d (%s)\n", MASK, ROWMASK, tpsm, per,
               148.0 * 4 * 32 * 1.965 / per, (int)same, cudaGetErrorString(e));
    }
}

int main() {
    int n = 148 * 2048;
    std::vector<fp_t> hx(n), hy(n);
    srand(1);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 8; ++k) {
            hx[i].l[k] = (uint32_t)rand() * 2654435761u + rand();
            hy[i].l[k] = (uint32_t)rand() * 40503u + rand();
            if (k == 7) { hx[i].l[k] &= 0x1fffffff; hy[i].l[k] &= 0x1fffffff; }
        }
    fp_t *dx, *dy;
    long long* d_cyc;
    cudaMalloc(&dx, n * sizeof(fp_t));
    cudaMalloc(&dy, n * sizeof(fp_t));
    cudaMalloc(&d_cyc, 4096 * sizeof(long long));
    cudaMemcpy(dy, hy.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
    run<0, 0xff>(dx, dy, hx, d_cyc);
    run<1, 0xff>(dx, dy, hx, d_cyc);
    run<2, 0xff>(dx, dy, hx, d_cyc);
    run<4, 0xff>(dx, dy, hx, d_cyc);
    run<8, 0xff>(dx, dy, hx, d_cyc);
    run<3, 0xff>(dx, dy, hx, d_cyc);
    run<12, 0xff>(dx, dy, hx, d_cyc);
    run<5, 0xff>(dx, dy, hx, d_cyc);
    run<10, 0xff>(dx, dy, hx, d_cyc);
    run<7, 0xff>(dx, dy, hx, d_cyc);
    run<15, 0xff>(dx, dy, hx, d_cyc);
    run<15, 0x55>(dx, dy, hx, d_cyc);
    run<5, 0x55>(dx, dy, hx, d_cyc);
    run<3, 0xaa>(dx, dy, hx, d_cyc);
    run<12, 0xaa>(dx, dy, hx, d_cyc);
    return 0;
}
