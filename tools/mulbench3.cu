// How does the 8 x 32-bit Montgomery multiplier of field.cuh scale with resident warps and with independent products per
// thread (ILP)?  Wall time (CUDA events over >= 100 ms runs) AND SM cycles (clock64) are both reported, so the
// cycles-per-multiplication figure does not depend on the clock the GPU happens to run at.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mulbench3 tools/mulbench3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../verkle_kzg_b200/csrc/field.cuh"
using namespace vk;

template <int ILP, bool NI>
__global__ void __launch_bounds__(128) kk(fp_t* x, const fp_t* y, int iters, long long* cyc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    fp_t a[ILP], b = y[i];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { a[j] = x[i]; a[j].l[0] += j; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = NI ? fp_mul_lazy_ni<FqParams>(a[j], b) : fp_mul_lazy<FqParams>(a[j], b);
    }
    long long t1 = clock64();
    fp_t r = a[0];
#pragma unroll
    for (int j = 1; j < ILP; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) r.l[k] ^= a[j].l[k];
    x[i] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, bool NI>
void run(fp_t* dx, fp_t* dy, long long* d_cyc) {
    int maxb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kk<ILP, NI>, 128, 0);
    for (int bps : {1, 2, 4, 8, 12, 16}) {   // blocks of 4 warps per SM = warps per SM sub-partition
        if (bps > maxb) continue;
        int blocks = 148 * bps;
        int iters = 40000 / ILP;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        kk<ILP, NI><<<blocks, 128>>>(dx, dy, 200, d_cyc);
        cudaEventRecord(e0);
        kk<ILP, NI><<<blocks, 128>>>(dx, dy, iters, d_cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(blocks);
        cudaMemcpy(h.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0, mx = 0;
        for (auto v : h) { avg += (double)v; if (v > mx) mx = (double)v; }
        avg /= blocks;
        double muls_per_warp = (double)iters * ILP;
        double per = avg / (muls_per_warp * bps);          // cycles per warp-multiplication per SM sub-partition
        double gmuls = (double)blocks * 128 * muls_per_warp / (ms * 1e-3) / 1e9;
        printf("ILP %d %s warps/SMSP %2d (max %2d): %7.1f cycles per warp-mul per SMSP | wall %8.2f ms -> %6.1f G mul/s | implied clock %5.0f MHz (%s)\n",
               ILP, NI ? "call  " : "inline", bps, maxb, per, ms, gmuls, mx / (ms * 1e-3) / 1e6, cudaGetErrorString(e));
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
    cudaMalloc(&d_cyc, 8192 * sizeof(long long));
    cudaMemcpy(dx, hx.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dy, hy.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
    run<1, false>(dx, dy, d_cyc);
    run<1, true>(dx, dy, d_cyc);
    run<2, false>(dx, dy, d_cyc);
    run<2, true>(dx, dy, d_cyc);
    run<4, false>(dx, dy, d_cyc);
    run<5, false>(dx, dy, d_cyc);
    return 0;
}
