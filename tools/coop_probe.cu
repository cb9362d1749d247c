// Latency of one point operation on a lone warp: plain (one lane, inlined products), lane pair (xyzz_add_pair), and the
// four-lane forms (one product per lane per dependency level, shuffle exchanges), the whole windowed multiplication, plus
// a bare chain of 10 multiplier calls for scale.  Results on B200: profiles/r01_coop_probe.txt.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/coop_probe tools/coop_probe.cu
#include <cstdio>
#include "../verkle_kzg_b200/csrc/warp_util.cuh"
using namespace vk;

// two / four independent products per call: does ptxas interleave the chains (instruction-level parallelism for a lone warp)?
static __device__ __noinline__ void fp_mul2_ni(fp_t& r0, fp_t& r1, const fp_t a0, const fp_t b0, const fp_t a1, const fp_t b1) {
    r0 = fp_mul<Q>(a0, b0);
    r1 = fp_mul<Q>(a1, b1);
}
static __device__ __noinline__ void fp_mul4_ni(fp_t& r0, fp_t& r1, fp_t& r2, fp_t& r3, const fp_t a0, const fp_t b0, const fp_t a1,
                                               const fp_t b1, const fp_t a2, const fp_t b2, const fp_t a3, const fp_t b3) {
    r0 = fp_mul<Q>(a0, b0);
    r1 = fp_mul<Q>(a1, b1);
    r2 = fp_mul<Q>(a2, b2);
    r3 = fp_mul<Q>(a3, b3);
}

__global__ void __launch_bounds__(32) k_probe(const xyzz_t* in, xyzz_t* out, long long* stamps) {
    const int lane = threadIdx.x;
    xyzz_t a = in[lane & 1], b = in[2 + (lane & 1)];
    affine_t P;
    P.x = in[1].x;
    P.y = in[1].y;
    long long t0 = clock64();
    fp_t m = a.x;
#pragma unroll 1
    for (int i = 0; i < 10; ++i) m = fp_mul_ni<Q>(m, b.y);
    long long t1 = clock64();
    xyzz_t r0 = xyzz_add_ni(a, b);
    long long t2 = clock64();
    xyzz_t r1 = xyzz_add_pair(a, 1);
    long long t3 = clock64();
    xyzz_t r2 = xyzz_add_quad(in[1], in[3]);
    long long t4 = clock64();
    xyzz_t r3 = xyzz_dbl_quad(in[0]);
    long long t5 = clock64();
    xyzz_t r4 = xyzz_madd_quad(in[0], P);
    long long t6 = clock64();
    xyzz_t r5 = xyzz_add_quad(in[0], in[2]);
    long long t7 = clock64();
    xyzz_t r6 = xyzz_dbl_ni(in[0]);
    long long t8 = clock64();
    long long u0 = clock64();
    fp_t d0 = a.x, d1 = a.y;
#pragma unroll 1
    for (int i = 0; i < 10; ++i) fp_mul2_ni(d0, d1, d0, b.y, d1, b.x);
    long long u1 = clock64();
    fp_t e0 = a.x, e1 = a.y, e2 = a.zz, e3 = a.zzz;
#pragma unroll 1
    for (int i = 0; i < 10; ++i) fp_mul4_ni(e0, e1, e2, e3, e0, b.y, e1, b.x, e2, b.zz, e3, b.zzz);
    long long u2 = clock64();
    r6.y = fp_add<Q>(fp_add<Q>(d0, d1), fp_add<Q>(fp_add<Q>(e0, e1), fp_add<Q>(e2, e3)));
    // the whole windowed multiplication: same scalar on every lane, then a different scalar per lane
    fp_t k = in[3].x;
    long long t9 = clock64();
    xyzz_t r7 = var_mul_windowed(P, k);
    long long t10 = clock64();
    k.l[0] += 0x9e3779b9u * lane; k.l[3] ^= 0x85ebca6bu * lane; k.l[6] += 0xc2b2ae35u * lane;
    xyzz_t r8 = var_mul_windowed(P, k);
    long long t11 = clock64();
    r6.x = fp_add<Q>(r6.x, fp_add<Q>(r7.x, r8.x));
    if (lane == 0) {
        stamps[8] = t10 - t9; stamps[9] = t11 - t10; stamps[10] = u1 - u0; stamps[11] = u2 - u1;
        stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = t3 - t2; stamps[3] = t4 - t3; stamps[4] = t5 - t4;
        stamps[5] = t6 - t5; stamps[6] = t7 - t6; stamps[7] = t8 - t7;
    }
    out[lane].y = r6.y;
    out[lane].x = fp_add<Q>(fp_add<Q>(fp_add<Q>(r0.x, r1.x), fp_add<Q>(r2.x, r3.x)), fp_add<Q>(fp_add<Q>(r4.x, r5.x), fp_add<Q>(r6.x, m)));
}

int main() {
    xyzz_t h[4];
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 8; ++i) {
            h[k].x.l[i] = 0x1234567u * (i + 1) + k;
            h[k].y.l[i] = 0x7654321u * (i + 3) + k;
            h[k].zz.l[i] = 0x1357911u * (i + 5) + k;
            h[k].zzz.l[i] = 0x2468aceu * (i + 7) + k;
        }
    for (int k = 0; k < 4; ++k) h[k].x.l[7] &= 0x0fffffff, h[k].y.l[7] &= 0x0fffffff, h[k].zz.l[7] &= 0x0fffffff, h[k].zzz.l[7] &= 0x0fffffff;
    xyzz_t *d_in, *d_out; long long* d_st;
    cudaMalloc(&d_in, sizeof(h)); cudaMalloc(&d_out, 32 * sizeof(xyzz_t)); cudaMalloc(&d_st, 128);
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    const char* names[12] = {"10 dependent fp_mul_ni", "xyzz_add_ni (1 lane, 14 products)", "xyzz_add_pair (2 lanes, 7 deep)",
                            "xyzz_add_quad (4 lanes, 4 deep), again", "xyzz_dbl_quad (3 deep)", "xyzz_madd_quad (4 deep)",
                            "xyzz_add_quad (4 deep)", "xyzz_dbl_ni (1 lane, 9 products)",
                            "var_mul_windowed, same scalar on all lanes", "var_mul_windowed, a scalar per lane",
                            "10 dependent calls of a 2-product multiplier", "10 dependent calls of a 4-product multiplier"};
    for (int rep = 0; rep < 3; ++rep) {
        k_probe<<<1, 32>>>(d_in, d_out, d_st);
        long long st[12];
        cudaMemcpy(st, d_st, 96, cudaMemcpyDeviceToHost);
        printf("run %d:\n", rep);
        for (int i = 0; i < 12; ++i) printf("  %-44s %8lld clk  %6.2f us\n", names[i], st[i], st[i] / 1965.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
