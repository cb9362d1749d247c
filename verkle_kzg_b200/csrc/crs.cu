// IPA CRS generation, the last "next" row of the scope table:
//   IPAPointGenerator::gen / gen_at        (vector-commit/src/ipa/ipa_point_generator.rs:51-81)
//   EthereumHashToCurve::hash              (:97-109)   SHA-256(seed || index as 8 LE bytes) -> Affine::from_random_bytes
// `gen(num)` is try-and-increment: the first `num` indices 0, 1, 2 ... whose digest parses as a curve point, in index
// order.  A candidate passes with probability ~0.19 (two flag bits valid 1/2, x < p 0.756, x^3 + 3 a square 1/2), so
// candidates are independent work — one thread each — and the only ordered step is the compaction of the survivors:
//   k_crs_candidates   thread per index: SHA-256 from the host-absorbed midstate of the seed's whole blocks, flag / range
//                      tests, y = rhs^((p+1)/4) (p = 3 mod 4; 252 squarings + 109 products, uniform control flow),
//                      y^2 == rhs test, root chosen by the flag bit
//   k_crs_compact      one CTA: ordered stream compaction (ballot + running offset) of the valid candidates into the
//                      first `need` output slots; records how many were found and the index after the last one used
// ark-ec 0.4 `from_random_bytes` (short Weierstrass, Fq::from_random_bytes_with_flags::<SWFlags>) restated: the root
// with the "larger" canonical y is taken when bit 7 of byte 31 is set — the same convention as affine_compress
// (curve.cuh), i.e. ONE constant decides both (tests/golden/arkworks `ipa_crs`).
#include "vk_common.cuh"
#include "hash.cuh"

namespace vk {

struct CrsSeed {
    uint32_t mid[8];    // SHA-256 state after the seed's whole 64-byte blocks
    uint8_t tail[64];   // the < 64 seed bytes that are left
    uint32_t tail_len;
    uint64_t seed_len;  // total seed length (for the length padding)
};

// (p + 1) / 4 for BN254 Fq, little-endian limbs
__device__ __constant__ uint32_t FQ_SQRT_EXP[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u,
                                                   0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};

__device__ __forceinline__ bool crs_candidate(const CrsSeed& sd, uint64_t index, affine_t& out) {
    sha256_ctx c;
#pragma unroll
    for (int i = 0; i < 8; ++i) c.h[i] = sd.mid[i];
    c.buflen = 0;
    c.total = sd.seed_len - sd.tail_len;
    sha256_update(c, sd.tail, sd.tail_len);
    uint8_t le[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) le[i] = (uint8_t)(index >> (8 * i));
    sha256_update(c, le, 8);
    uint8_t d[32];
    sha256_final(c, d);

    const uint32_t flags = d[31] & 0xC0u;
    d[31] &= 0x3Fu;
    fp_t x;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        x.l[i] = (uint32_t)d[4 * i] | ((uint32_t)d[4 * i + 1] << 8) | ((uint32_t)d[4 * i + 2] << 16) | ((uint32_t)d[4 * i + 3] << 24);
    uint32_t pl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pl[i] = Q::p(i);
    if (geq8(x.l, pl)) return false;   // Fq::deserialize of the masked bytes fails
    if (flags == 0xC0u) return false;  // SWFlags::from_u8: both bits set
    if (flags == 0x40u) {              // infinity flag: the identity iff x == 0
        if (!fp_is_zero(x)) return false;
        out.x = fp_zero<Q>();
        out.y = fp_zero<Q>();
        return true;
    }
    const fp_t xm = fp_to_mont<Q>(x);
    fp_t three = fp_one<Q>();
    three = fp_add<Q>(fp_add<Q>(three, three), fp_one<Q>());
    const fp_t rhs = fp_add<Q>(fp_mul_ni<Q>(fp_mul_ni<Q>(xm, xm), xm), three);
    fp_t y = fp_one<Q>();
#pragma unroll 1
    for (int bit = 251; bit >= 0; --bit) {  // the exponent has 252 bits
        y = fp_mul_ni<Q>(y, y);
        if ((FQ_SQRT_EXP[bit >> 5] >> (bit & 31)) & 1) y = fp_mul_ni<Q>(y, rhs);
    }
    if (!fp_eq(fp_mul_ni<Q>(y, y), rhs)) return false;  // non-residue: no point with this x
    const bool want_larger = flags == 0x80u;
    if (fp_is_lexicographically_largest<Q>(y) != want_larger) y = fp_neg<Q>(y);
    out.x = xm;
    out.y = y;
    return true;
}

__global__ void __launch_bounds__(128) k_crs_candidates(CrsSeed sd, uint64_t first, uint64_t count, affine_t* __restrict__ pts,
                                                        uint8_t* __restrict__ valid) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    affine_t p;
    const bool ok = crs_candidate(sd, first + i, p);
    valid[i] = ok ? 1 : 0;
    if (ok) {
        fp_store(&pts[i].x, p.x);
        fp_store(&pts[i].y, p.y);
    }
}

// state[0] = points written so far, state[1] = index after the candidate that completed the request (0 = not yet)
__global__ void __launch_bounds__(1024) k_crs_compact(const affine_t* __restrict__ pts, const uint8_t* __restrict__ valid, uint64_t first,
                                                      uint64_t count, uint64_t need, affine_t* __restrict__ out, uint64_t* __restrict__ state) {
    __shared__ uint32_t warp_cnt[32];
    __shared__ uint64_t base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = state[0];
    __syncthreads();
    for (uint64_t start = 0; start < count; start += blockDim.x) {
        const uint64_t i = start + threadIdx.x;
        const bool v = i < count && valid[i];
        const uint32_t bal = __ballot_sync(0xffffffffu, v);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w) {
            const uint32_t c = warp_cnt[w];
            if (w < warp) before += c;
            total += c;
        }
        const uint64_t slot = base + before + __popc(bal & ((1u << lane) - 1u));
        if (v && slot < need) {
            const uint4* src = reinterpret_cast<const uint4*>(pts + i);
            uint4* dst = reinterpret_cast<uint4*>(out + slot);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = src[q];
            if (slot + 1 == need) state[1] = first + i + 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
        if (base >= need) break;
    }
    if (threadIdx.x == 0) state[0] = base < need ? base : need;
}

static void crs_seed(CrsSeed& sd, const uint8_t* seed, uint64_t seed_len) {
    const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memset(&sd, 0, sizeof(sd));
    for (int i = 0; i < 8; ++i) sd.mid[i] = IV[i];
    const uint64_t nblk = seed_len / 64;
    for (uint64_t b = 0; b < nblk; ++b) sha256_compress(sd.mid, seed + 64 * b);
    sd.tail_len = (uint32_t)(seed_len - 64 * nblk);
    if (sd.tail_len) memcpy(sd.tail, seed + 64 * nblk, sd.tail_len);
    sd.seed_len = seed_len;
}

}  // namespace vk

using namespace vk;

extern "C" {

int32_t vkzg_ipa_crs_generate(vkzg_ctx* ctx, const uint8_t* seed, uint64_t seed_len, uint64_t num, vkzg_g1_affine* out,
                              uint64_t* next_index) {
    VK_TRY(ctx_check(ctx));
    if ((seed_len && !seed) || !num || !out) return VKZG_ERR_ARG;
    if (num > (1ull << 28)) return VKZG_ERR_RANGE;
    CrsSeed sd;
    crs_seed(sd, seed, seed_len);
    DevBuf<affine_t> d_out, d_pts;
    DevBuf<uint8_t> d_valid;
    DevBuf<uint64_t> d_state;
    VK_TRY(d_out.alloc(ctx, num));
    VK_TRY(d_state.alloc(ctx, 2));
    VK_CUDA(cudaMemsetAsync(d_state.p, 0, 2 * sizeof(uint64_t), ctx->stream));
    // a candidate survives with probability ~0.189: the first pass covers the mean + a margin, later passes the shortfall
    const uint64_t cap = 1ull << 22;
    uint64_t first = 0, have = 0;
    uint64_t chunk = num * 11 / 2 + 256;
    if (chunk > cap) chunk = cap;
    VK_TRY(d_pts.alloc(ctx, chunk));
    VK_TRY(d_valid.alloc(ctx, chunk));
    uint64_t state[2] = {0, 0};
    while (have < num) {
        uint64_t count = (num - have) * 11 / 2 + 256;
        if (count > chunk) count = chunk;
        k_crs_candidates<<<ceil_div_u64(count, 128), 128, 0, ctx->stream>>>(sd, first, count, d_pts, d_valid);
        VK_TRY(launch_check(ctx));
        k_crs_compact<<<1, 1024, 0, ctx->stream>>>(d_pts, d_valid, first, count, num, d_out, d_state);
        VK_TRY(launch_check(ctx));
        VK_TRY(download(ctx, state, d_state.p, 2));
        VK_TRY(stream_sync(ctx));
        have = state[0];
        first += count;
    }
    VK_TRY(download(ctx, out, d_out.p, num));
    VK_TRY(stream_sync(ctx));
    if (next_index) *next_index = state[1];
    return VKZG_OK;
}

int32_t vkzg_ipa_crs_generate_at(vkzg_ctx* ctx, const uint8_t* seed, uint64_t seed_len, uint64_t index, vkzg_g1_affine* out,
                                 int32_t* ok) {
    VK_TRY(ctx_check(ctx));
    if ((seed_len && !seed) || !out || !ok) return VKZG_ERR_ARG;
    CrsSeed sd;
    crs_seed(sd, seed, seed_len);
    DevBuf<affine_t> d_pts;
    DevBuf<uint8_t> d_valid;
    VK_TRY(d_pts.alloc(ctx, 1));
    VK_TRY(d_valid.alloc(ctx, 1));
    VK_CUDA(cudaMemsetAsync(d_pts.p, 0, sizeof(affine_t), ctx->stream));
    k_crs_candidates<<<1, 128, 0, ctx->stream>>>(sd, index, 1, d_pts, d_valid);
    VK_TRY(launch_check(ctx));
    uint8_t v = 0;
    VK_TRY(download(ctx, &v, d_valid.p, 1));
    VK_TRY(download(ctx, out, d_pts.p, 1));
    VK_TRY(stream_sync(ctx));
    *ok = v ? 1 : 0;
    if (!v) memset(out, 0, sizeof(*out));
    return VKZG_OK;
}

}  // extern "C"
