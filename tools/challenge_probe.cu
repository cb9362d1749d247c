// Where does one thread's round tail of an IPA proof (k_ipa_challenge, ipa.cu) spend its time?  One lane runs the same
// sequence — shared inversion, two normalisations, transcript appends, digest — with clock64() stamps between the phases.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/challenge_probe tools/challenge_probe.cu
#include <cstdio>
#include "../verkle_kzg_b200/csrc/warp_util.cuh"
using namespace vk;

__global__ void __launch_bounds__(64) k_probe(const xyzz_t* lr, transcript_t* tr, fp_t* x_out, affine_t* out, long long* stamps) {
    const bool live = threadIdx.x == 0;
    long long t0 = clock64();
    xyzz_t l = lr[0], r = lr[1];
    fp_t zl = l.zzz, zr = r.zzz;
    long long t1 = clock64();
    fp_t inv = warp_inverse_of_lane_products_t<Q>(live ? fp_mul_ni<Q>(zl, zr) : fp_one<Q>());
    long long t2 = clock64();
    fp_t inv_direct = fp_inv<Q>(fp_mul_ni<Q>(zl, zr));
    long long t3 = clock64();
    if (!live) return;
    affine_t la = xyzz_to_affine_with_inv(l, fp_mul_ni<Q>(inv, zr));
    affine_t ra = xyzz_to_affine_with_inv(r, fp_mul_ni<Q>(inv_direct, zl));
    out[0] = la;
    out[1] = ra;
    long long t4 = clock64();
    transcript_t t = tr[0];
    long long t5 = clock64();
    tr_append_point(t, la, "L");
    tr_append_point(t, ra, "R");
    long long t6 = clock64();
    fp_t x = tr_digest(t, "x");
    long long t7 = clock64();
    fp_store(x_out, x);
    tr[0] = t;
    long long t8 = clock64();
    // the digest's pieces: five bare compressions of a register block, and the XMD without the reduction
    uint32_t hh[8], ww[16];
    for (int i = 0; i < 8; ++i) hh[i] = x.l[i];
    for (int i = 0; i < 16; ++i) ww[i] = x.l[i & 7] + i;
    for (int k = 0; k < 5; ++k) sha256_compress_words(hh, ww);
    long long t9 = clock64();
    uint32_t b1[8], b2[8];
    xmd48_words(t.state, 100, t.dst, t.dst_len, 48, b1, b2);
    long long t10 = clock64();
    fp_t red = fr_from_be48_words(b1, b2);
    long long t11 = clock64();
    x_out[1].l[0] = hh[0] ^ red.l[0];
    stamps[8] = t9 - t8; stamps[9] = t10 - t9; stamps[10] = t11 - t10;
    stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = t3 - t2; stamps[3] = t4 - t3; stamps[4] = t5 - t4;
    stamps[5] = t6 - t5; stamps[6] = t7 - t6; stamps[7] = t8 - t7;
}

int main() {
    xyzz_t h[2];
    // two arbitrary non-trivial XYZZ points: (x, y, zz, zzz) need not be on the curve for timing purposes
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < 8; ++i) {
            h[k].x.l[i] = 0x1234567u * (i + 1) + k;
            h[k].y.l[i] = 0x7654321u * (i + 3) + k;
            h[k].zz.l[i] = 0x1357911u * (i + 5) + k;
            h[k].zzz.l[i] = 0x2468aceu * (i + 7) + k;
        }
    for (int k = 0; k < 2; ++k) h[k].x.l[7] &= 0x0fffffff, h[k].y.l[7] &= 0x0fffffff, h[k].zz.l[7] &= 0x0fffffff, h[k].zzz.l[7] &= 0x0fffffff;
    xyzz_t* d_lr; transcript_t* d_tr; fp_t* d_x; affine_t* d_out; long long* d_st;
    cudaMalloc(&d_lr, sizeof(h)); cudaMalloc(&d_tr, sizeof(transcript_t)); cudaMalloc(&d_x, 64); cudaMalloc(&d_out, 128); cudaMalloc(&d_st, 128);
    cudaMemcpy(d_lr, h, sizeof(h), cudaMemcpyHostToDevice);
    transcript_t t{};
    t.len = 33; t.dst[0] = 'i'; t.dst[1] = 'p'; t.dst[2] = 'a'; t.dst_len = 3;
    const char* names[11] = {"load L,R", "warp-shared inversion (scan + Kaliski)", "direct Kaliski inversion", "2 x to-affine + store",
                            "load transcript", "serialise + append L, R", "digest (XMD + reduce)", "store",
                            "5 bare compressions", "xmd48_words(100 bytes)", "fr_from_be48_words"};
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(d_tr, &t, sizeof(t), cudaMemcpyHostToDevice);
        k_probe<<<1, 64>>>(d_lr, d_tr, d_x, d_out, d_st);
        long long st[11];
        cudaMemcpy(st, d_st, 88, cudaMemcpyDeviceToHost);
        printf("run %d (%s):\n", rep, rep ? "warm" : "cold");
        long long tot = 0;
        for (int i = 0; i < 11; ++i) {
            printf("  %-42s %8lld clk  %6.1f us\n", names[i], st[i], st[i] / 1965.0);
            if (i < 8) tot += st[i];
        }
        printf("  total %lld clk = %.1f us\n", tot, tot / 1965.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
