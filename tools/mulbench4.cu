// Multiplier throughput against RESIDENT WARPS, measured so that the answer cannot depend on how the block scheduler
// spreads an under-filled grid: dynamic shared memory limits every SM to exactly `bps` blocks of 128 threads (one warp
// per SM sub-partition each), the grid is exactly 148 * bps blocks (one full wave, every SM equally loaded — checked
// through %smid), and the figure is WALL time of a >= 50 ms run.  Variants: the library's 8 x 32-bit Montgomery product
// inline / called; 1 or 2 independent products per thread; the XYZZ mixed addition of the hot loops.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mulbench4 tools/mulbench4.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../verkle_kzg_b200/csrc/curve.cuh"
using namespace vk;

extern __shared__ uint32_t dyn_smem[];

template <int MODE>  // 0: inline mul, 1: called mul, 2: two independent inline muls, 3: xyzz_madd_hot
__global__ void __launch_bounds__(128) kk(fp_t* x, const fp_t* y, int iters, uint32_t* smid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) {
        uint32_t s;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        smid[blockIdx.x] = s;
        dyn_smem[0] = s;
    }
    fp_t a = x[i], b = y[i];
    if (MODE == 0 || MODE == 1) {
#pragma unroll 1
        for (int k = 0; k < iters; ++k) a = MODE == 0 ? fp_mul_lazy<FqParams>(a, b) : fp_mul_lazy_ni<FqParams>(a, b);
        x[i] = a;
    } else if (MODE == 2) {
        fp_t a2 = y[i];
        a2.l[0] ^= 5;
#pragma unroll 1
        for (int k = 0; k < iters; ++k) {
            a = fp_mul_lazy<FqParams>(a, b);
            a2 = fp_mul_lazy<FqParams>(a2, b);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) a.l[k] ^= a2.l[k];
        x[i] = a;
    } else {
        xyzz_t acc;
        acc.x = a;
        acc.y = b;
        acc.zz = a;
        acc.zzz = b;
        affine_t p;
        p.x = b;
        p.y = a;
#pragma unroll 1
        for (int k = 0; k < iters; ++k) {
            xyzz_madd_hot(acc, p);
            p.x.l[0] ^= (uint32_t)k;  // (keeps the point changing; not a curve point — the products do not care)
        }
        x[i] = acc.x;
    }
}

template <int MODE>
void run(const char* name, fp_t* dx, fp_t* dy, uint32_t* d_smid, int muls_per_iter) {
    cudaFuncSetAttribute(kk<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kk<MODE>);
    for (int bps : {1, 2, 3, 4, 5, 6, 8, 10, 12, 16}) {
        size_t smem = (size_t)(220 * 1024) / bps;
        smem &= ~(size_t)255;
        int maxb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kk<MODE>, 128, smem);
        if (maxb != bps) {
            printf("%-28s warps/SMSP %2d: occupancy gives %d blocks/SM (regs %d) — skipped\n", name, bps, maxb, fa.numRegs);
            continue;
        }
        int blocks = 148 * bps;
        int iters = 24000 / muls_per_iter * (MODE == 3 ? 1 : 1);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        kk<MODE><<<blocks, 128, smem>>>(dx, dy, 100, d_smid);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        kk<MODE><<<blocks, 128, smem>>>(dx, dy, iters, d_smid);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<uint32_t> h(blocks);
        cudaMemcpy(h.data(), d_smid, blocks * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        std::vector<int> cnt(256, 0);
        for (auto s : h) cnt[s & 255]++;
        int mn = 1 << 30, mx = 0, used = 0;
        for (int s = 0; s < 256; ++s)
            if (cnt[s]) {
                ++used;
                mn = cnt[s] < mn ? cnt[s] : mn;
                mx = cnt[s] > mx ? cnt[s] : mx;
            }
        double muls = (double)blocks * 128 * iters * muls_per_iter;
        double gm = muls / (ms * 1e-3) / 1e9;
        printf("%-28s warps/SMSP %2d regs %3d: %7.2f ms  %6.1f G mul/s  = %5.2f T MAC32/s | blocks per SM min %d max %d on %d SMs (%s)\n", name, bps,
               fa.numRegs, ms, gm, gm * 136 / 1e3, mn, mx, used, cudaGetErrorString(e));
    }
}

int main() {
    int n = 148 * 16 * 128;
    std::vector<fp_t> hx(n), hy(n);
    srand(1);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 8; ++k) {
            hx[i].l[k] = (uint32_t)rand() * 2654435761u + rand();
            hy[i].l[k] = (uint32_t)rand() * 40503u + rand();
            if (k == 7) { hx[i].l[k] &= 0x1fffffff; hy[i].l[k] &= 0x1fffffff; }
        }
    fp_t *dx, *dy;
    uint32_t* d_smid;
    cudaMalloc(&dx, n * sizeof(fp_t));
    cudaMalloc(&dy, n * sizeof(fp_t));
    cudaMalloc(&d_smid, 148 * 16 * sizeof(uint32_t));
    cudaMemcpy(dx, hx.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dy, hy.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
    run<0>("fp_mul_lazy inline", dx, dy, d_smid, 1);
    run<1>("fp_mul_lazy called", dx, dy, d_smid, 1);
    run<2>("2 independent inline", dx, dy, d_smid, 2);
    run<3>("xyzz_madd_hot (10 mul)", dx, dy, d_smid, 10);
    return 0;
}
