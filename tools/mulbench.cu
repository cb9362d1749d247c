// Microbenchmark: Fq multiplier variants on sm_100a (throughput of dependent chains, many threads).
//   variant 0: 8 x 32-bit limbs, mad.lo.cc/madc.hi.cc chains (field.cuh fp_mul)
//   variant 1: 9 x 29-bit limbs, carry-free mad.wide.u32 column accumulation (R' = 2^261)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mulbench tools/mulbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../verkle_kzg_b200/csrc/field.cuh"
using namespace vk;

struct fq29 { uint32_t l[9]; };
static const uint32_t M29 = (1u << 29) - 1;

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
__host__ __device__ inline fp_t from29(const fq29& a) {  // limbs must be normalised and value < 2^256
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

struct P29 {
    uint32_t p[9];
    uint32_t inv;  // -p^-1 mod 2^29
};
__constant__ P29 c_p29;

__device__ __forceinline__ fq29 mul29(const fq29& a, const fq29& b) {
    uint64_t t[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) t[i + j] += (uint64_t)a.l[j] * b.l[i];
        uint32_t m = ((uint32_t)t[i] * c_p29.inv) & M29;
#pragma unroll
        for (int j = 0; j < 9; ++j) t[i + j] += (uint64_t)m * c_p29.p[j];
        t[i + 1] += t[i] >> 29;
    }
    fq29 r;
    uint64_t c = t[9];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        r.l[k] = (uint32_t)c & M29;
        c = (c >> 29) + t[10 + k];
    }
    r.l[8] = (uint32_t)c;
    return r;
}

__global__ void k0(fp_t* x, const fp_t* y, int n, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t a = x[i], b = y[i];
#pragma unroll 1
    for (int k = 0; k < iters; ++k) a = fp_mul<FqParams>(a, b);
    x[i] = a;
}
__global__ void k1(fp_t* x, const fp_t* y, int n, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq29 a = to29(x[i]), b = to29(y[i]);
#pragma unroll 1
    for (int k = 0; k < iters; ++k) a = mul29(a, b);
    // normalise fully for output: value < 2p (< 2^255) fits 256 bits
    x[i] = from29(a);
}

int main(int argc, char** argv) {
    int sms = 148;
    // p in 29-bit limbs, inv
    P29 h;
    fp_t pp;
    for (int i = 0; i < 8; ++i) pp.l[i] = FqParams::p(i);
    fq29 p29 = to29(pp);
    for (int i = 0; i < 9; ++i) h.p[i] = p29.l[i];
    // inv = -p^-1 mod 2^29 via Newton
    uint32_t p0 = h.p[0], inv = 1;
    for (int i = 0; i < 6; ++i) inv *= 2 - p0 * inv;
    h.inv = (0u - inv) & M29;
    cudaMemcpyToSymbol(c_p29, &h, sizeof(h));
    for (int tpsm : {256, 512, 1024, 2048}) {
        int n = sms * tpsm;
        std::vector<fp_t> hx(n), hy(n);
        srand(1);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 8; ++k) {
                hx[i].l[k] = (uint32_t)rand() * 2654435761u + rand();
                hy[i].l[k] = (uint32_t)rand() * 40503u + rand();
                if (k == 7) { hx[i].l[k] &= 0x1fffffff; hy[i].l[k] &= 0x1fffffff; }
            }
        fp_t *dx, *dy;
        cudaMalloc(&dx, n * sizeof(fp_t));
        cudaMalloc(&dy, n * sizeof(fp_t));
        for (int variant = 0; variant < 2; ++variant) {
            cudaMemcpy(dx, hx.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
            cudaMemcpy(dy, hy.data(), n * sizeof(fp_t), cudaMemcpyHostToDevice);
            int iters = 2000;
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(a);
                if (variant == 0) k0<<<n / 256, 256>>>(dx, dy, n, iters); else k1<<<n / 256, 256>>>(dx, dy, n, iters);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
            }
            printf("variant %d threads/SM %4d : %8.2f Gmul/s  (err %s)\n", variant, tpsm, (double)n * iters / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        if (tpsm == 256) {
            // correctness dump: one multiplication per variant on the first 4 elements
            FILE* f = fopen("gpurun_out/mulbench.txt", "w");
            for (int variant = 0; variant < 2; ++variant) {
                cudaMemcpy(dx, hx.data(), 4 * sizeof(fp_t), cudaMemcpyHostToDevice);
                cudaMemcpy(dy, hy.data(), 4 * sizeof(fp_t), cudaMemcpyHostToDevice);
                if (variant == 0) k0<<<1, 256>>>(dx, dy, 4, 1); else k1<<<1, 256>>>(dx, dy, 4, 1);
                std::vector<fp_t> out(4);
                cudaMemcpy(out.data(), dx, 4 * sizeof(fp_t), cudaMemcpyDeviceToHost);
                for (int i = 0; i < 4; ++i) {
                    fprintf(f, "v%d ", variant);
                    for (int k = 7; k >= 0; --k) fprintf(f, "%08x", hx[i].l[k]);
                    fprintf(f, " ");
                    for (int k = 7; k >= 0; --k) fprintf(f, "%08x", hy[i].l[k]);
                    fprintf(f, " ");
                    for (int k = 7; k >= 0; --k) fprintf(f, "%08x", out[i].l[k]);
                    fprintf(f, "\n");
                }
            }
            fclose(f);
        }
        cudaFree(dx);
        cudaFree(dy);
    }
    return 0;
}
