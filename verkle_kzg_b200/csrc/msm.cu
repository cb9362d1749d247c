// M1 at large n (BASELINE config 4): one Pippenger MSM over a VKZG_KEY_MSM key.
//
// The key holds T[w][i] = 2^(c w) * P_i, so every signed c-bit digit of every scalar lands in ONE shared
// set of 2^(c-1) buckets (no per-window bucket sets, no final doubling chain):
//     sum_i s_i P_i = sum_b (b + 1) * B_b ,   B_b = sum over digits of magnitude b + 1 of +-T[w][i].
//   1. k_msm_scatter_fixed   one thread per scalar: Montgomery -> canonical, signed recode, every digit appended to the
//                     fixed-capacity list of its bucket (one pass; an overflow is counted on the device and the exact
//                     counting sort — k_msm_digits x2 + k_scan — enqueued behind it runs only then)
//   2. k_msm_order    buckets in order of decreasing list length (the lanes of a warp walk lists of equal length)
//   3. k_msm_bucket   P lanes per bucket (P adjacent lanes, interleaved slices -> equal work), 64-byte
//                     vectorised gathers one entry ahead of the mixed addition, shuffle-tree fold of the
//                     P partial sums
//   4. weighted bucket sum: k_msm_bitsums + k_msm_bitcombine (bit-parallel) up to 4096 buckets, above that k_msm_group_sums
//      (h + l plain row / column sums) followed by the same two kernels over h + l elements; then the common normalisation
// The order in which a bucket's points are added is not deterministic (atomics), the result is: it leaves
// the device in canonical affine form.
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

__device__ __forceinline__ affine_t load_affine_ro2(const affine_t* p) {
    affine_t a;
    a.x = fp_load_ro(&p->x);
    a.y = fp_load_ro(&p->y);
    return a;
}

// `gate` (may be null): the kernels of the exact two-pass path run only when the optimistic single pass dropped an
// entry (*gate != 0), the optimistic bucket pass only when it did not — decided ON THE DEVICE, so that an MSM call only
// enqueues work (no host round trip, capturable in a CUDA graph).
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const fp_t* __restrict__ scalars, uint64_t n, uint32_t c, uint32_t W,
                                                    uint32_t key_n, uint64_t first, uint32_t* __restrict__ hist_or_cursor,
                                                    const uint32_t* __restrict__ offsets, uint32_t* __restrict__ entries,
                                                    const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t k = fp_from_mont<S>(fp_load_ro(scalars + i));
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < W; ++w) {
        uint32_t v = scalar_bits(k.l, w * c, c) + carry;
        uint32_t neg = v >= half && w + 1 < W ? 1u : 0u;
        uint32_t mag = neg ? (1u << c) - v : v;
        carry = neg;
        if (mag) {
            uint32_t b = mag - 1;
            uint32_t pos = atomicAdd(hist_or_cursor + b, 1u);
            if (SCATTER) entries[offsets[b] + pos] = (uint32_t)((uint64_t)w * key_n + first + i) | (neg << 31);
        }
    }
}

// Optimistic single pass: every bucket owns `cap` slots (mean size + 25 %), entries that do not fit are counted in
// *dropped and the caller falls back to the exact two-pass counting sort (skewed scalars).
// The TOP window holds only 254 - c (W - 1) bits, i.e. every scalar's top digit falls into the first `top_n` buckets
// (49 of them at c = 13): those buckets get `top_extra` more slots each — list b starts at b * cap + min(b, top_n) * top_extra.
__device__ __forceinline__ size_t msm_list_base(uint32_t b, uint32_t cap, uint32_t top_n, uint32_t top_extra) {
    return (size_t)b * cap + (size_t)(b < top_n ? b : top_n) * top_extra;
}
__global__ void __launch_bounds__(256) k_msm_scatter_fixed(const fp_t* __restrict__ scalars, uint64_t n, uint32_t c, uint32_t W,
                                                           uint32_t key_n, uint64_t first, uint32_t cap, uint32_t top_n, uint32_t top_extra,
                                                           uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries,
                                                           uint32_t* __restrict__ dropped) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t k = fp_from_mont<S>(fp_load_ro(scalars + i));
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0, lost = 0;
    // four digits at a time: their atomics are issued back to back (four cursor round trips in flight per thread), then the
    // four entry stores (measured at 2^20 points: see DESIGN.md, MSM flow)
    for (uint32_t w0 = 0; w0 < W; w0 += 4) {
        uint32_t b[4], ent[4], pos[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t w = w0 + j;
            b[j] = 0xffffffffu;
            if (w < W) {
                uint32_t v = scalar_bits(k.l, w * c, c) + carry;
                uint32_t neg = v >= half && w + 1 < W ? 1u : 0u;
                uint32_t mag = neg ? (1u << c) - v : v;
                carry = neg;
                if (mag) {
                    b[j] = mag - 1;
                    ent[j] = (uint32_t)((uint64_t)w * key_n + first + i) | (neg << 31);
                }
            }
        }
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (b[j] != 0xffffffffu) pos[j] = atomicAdd(cursor + b[j], 1u);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (b[j] != 0xffffffffu) {
                if (pos[j] < cap + (b[j] < top_n ? top_extra : 0u))
                    entries[msm_list_base(b[j], cap, top_n, top_extra) + pos[j]] = ent[j];
                else
                    ++lost;
            }
    }
    if (lost) atomicAdd(dropped, lost);
}

// single-CTA exclusive scan of `n` counts (n <= 2^20), also zeroes the counts for their second life as cursors
__global__ void __launch_bounds__(1024) k_scan(uint32_t* __restrict__ counts, uint32_t n, uint32_t* __restrict__ offsets,
                                               const uint32_t* __restrict__ gate) {
    if (gate && *gate == 0) return;
    __shared__ uint32_t sh[1024];
    uint32_t per = (n + 1023) / 1024;
    uint32_t lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += counts[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = sh[threadIdx.x] - s;
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t cnt = counts[i];
        offsets[i] = run;
        counts[i] = 0;
        run += cnt;
    }
    if (threadIdx.x == 1023) offsets[n] = sh[1023];
}

// fallback taken: the bucket sizes of the optimistic pass become the zeroed histogram of the counting sort
__global__ void __launch_bounds__(256) k_msm_clear_if(uint32_t* __restrict__ counts, uint32_t n, const uint32_t* __restrict__ gate) {
    if (*gate == 0) return;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = 0;
}

// Buckets in order of decreasing list length (single CTA: shared-memory histogram of the sizes, scan, scatter).  The lanes of
// a warp serve 32 / P buckets and wait for the longest of their lists (sizes ~ Poisson: +-2 additions on 30 per lane at
// n = 2^20, c = 17); with neighbours of equal size that wait disappears, and the longest lists start first.
static const int ORD_BINS = 2048;
__global__ void __launch_bounds__(1024) k_msm_order(const uint32_t* __restrict__ sizes, uint32_t nb, uint32_t* __restrict__ order) {
    __shared__ uint32_t hist[ORD_BINS];
    __shared__ uint32_t part[1024];
    for (uint32_t i = threadIdx.x; i < ORD_BINS; i += 1024) hist[i] = 0;
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += 1024) {
        uint32_t sz = sizes[b];
        atomicAdd(&hist[ORD_BINS - 1 - (sz < ORD_BINS ? sz : ORD_BINS - 1)], 1u);
    }
    __syncthreads();
    const uint32_t h0 = hist[2 * threadIdx.x], h1 = hist[2 * threadIdx.x + 1];
    part[threadIdx.x] = h0 + h1;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const uint32_t base = part[threadIdx.x] - (h0 + h1);
    hist[2 * threadIdx.x] = base;
    hist[2 * threadIdx.x + 1] = base + h0;
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += 1024) {
        uint32_t sz = sizes[b];
        order[atomicAdd(&hist[ORD_BINS - 1 - (sz < ORD_BINS ? sz : ORD_BINS - 1)], 1u)] = b;
    }
}

// P (power of two <= 32) adjacent lanes per bucket
__global__ void __launch_bounds__(128, 4) k_msm_bucket(const affine_t* __restrict__ table, const uint32_t* __restrict__ offsets,
                                                    const uint32_t* __restrict__ entries, uint32_t nb, uint32_t P, uint32_t cap,
                                                    uint32_t top_n, uint32_t top_extra, xyzz_t* __restrict__ buckets,
                                                    const uint32_t* __restrict__ gate, uint32_t gate_want,
                                                    const uint32_t* __restrict__ order) {
    if (gate && (*gate != 0) != (gate_want != 0)) return;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = (uint32_t)(t / P), p = (uint32_t)(t % P);
    bool live = b < nb;
    if (live && order) b = __ldg(order + b);
    size_t lo = 0, hi = 0;
    if (live) {
        if (cap) {  // fixed-capacity layout: offsets[] holds the bucket sizes
            lo = msm_list_base(b, cap, top_n, top_extra);
            hi = lo + offsets[b];
        } else {
            lo = offsets[b];
            hi = offsets[b + 1];
        }
    }
    xyzz_t acc = xyzz_inf();
    size_t i = lo + p;
    uint32_t e = 0;
    affine_t cur;
    if (i < hi) {
        e = __ldg(entries + i);
        cur = load_affine_ro2(table + (e & 0x7fffffffu));
    }
    while (i < hi) {
        size_t in = i + P;
        uint32_t en = 0;
        affine_t nxt;
        if (in < hi) {
            en = __ldg(entries + in);
            nxt = load_affine_ro2(table + (en & 0x7fffffffu));
        }
        if (e >> 31) cur.y = fp_neg<Q>(cur.y);
        xyzz_madd_hot(acc, cur);
        cur = nxt;
        e = en;
        i = in;
    }
    xyzz_canon(acc);
#pragma unroll 1
    for (uint32_t m = 1; m < P; m <<= 1) acc = xyzz_add_pair(acc, (int)m);
    if (live && p == 0) {
        fp_store(&buckets[b].x, acc.x);
        fp_store(&buckets[b].y, acc.y);
        fp_store(&buckets[b].zz, acc.zz);
        fp_store(&buckets[b].zzz, acc.zzz);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weighted bucket sum  sum_b (b + 1) B_b.  This tail is a chain of DEPENDENT point operations (each ~10 k cycles for a
// lone warp, whatever the GPU has idle), so it is organised for depth, not for work: with w = b + 1 written in binary,
//     sum_b w_b B_b = sum_j 2^j T_j ,   T_j = sum of the buckets whose weight has bit j set,
// every T_j is a PLAIN sum (tree, no running sums, no per-segment scalar multiple): k_msm_bitsums — one CTA per
// (bit, slice of 2048 buckets): <= 4 additions per thread + a pair-split shuffle tree + 3 shared-memory levels;
// k_msm_bitcombine — warp j folds the slice sums of bit j, doubles j times, the c values are summed.  ~35 dependent
// operations, most of them doublings or pair-split additions, against the ~52 full additions of the running-sum form
// (measured at n = 2^16: 325 -> see profiles/ us for the two kernels).
// ---------------------------------------------------------------------------------------------------------------
static const int BS_THREADS = 128;
static const int BS_SLICE = 512;   // buckets per CTA (4 warps: one per sub-partition, so the dependent additions do not share a multiplier pipe)

// weight of element e: plain bucket sets weigh b + 1; the group sums of the two-level form (split_h != 0) weigh
// e * l for the h per-group sums S_e (e < split_h) and (e - split_h) + 1 for the l per-position sums T_(e - split_h)
__device__ __forceinline__ uint32_t msm_weight(uint32_t e, uint32_t split_h, uint32_t l) {
    return split_h == 0 ? e + 1 : (e < split_h ? e * l : e - split_h + 1);
}

__global__ void __launch_bounds__(BS_THREADS) k_msm_bitsums(const xyzz_t* __restrict__ buckets, uint32_t nb, uint32_t slices,
                                                            uint32_t split_h, uint32_t split_l,
                                                            xyzz_t* __restrict__ out /*[bits][slices]*/) {
    __shared__ xyzz_t sh[BS_THREADS / 32];
    const uint32_t bit = blockIdx.y, slice = blockIdx.x;
    const uint32_t lo = slice * BS_SLICE;
    xyzz_t acc = xyzz_inf();
#pragma unroll 1
    for (uint32_t i = threadIdx.x; i < BS_SLICE; i += BS_THREADS) {
        uint32_t b = lo + i;
        if (b < nb && ((msm_weight(b, split_h, split_l) >> bit) & 1)) {
            xyzz_t B;
            B.x = fp_load(&buckets[b].x);
            B.y = fp_load(&buckets[b].y);
            B.zz = fp_load(&buckets[b].zz);
            B.zzz = fp_load(&buckets[b].zzz);
            acc = xyzz_add_ni(acc, B);
        }
    }
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) acc = xyzz_add_pair(acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        xyzz_t v = threadIdx.x < BS_THREADS / 32 ? sh[threadIdx.x] : xyzz_inf();
#pragma unroll 1
        for (int off = 1; off < BS_THREADS / 32; off <<= 1) v = xyzz_add_pair(v, off);
        if (threadIdx.x == 0) out[bit * slices + slice] = v;
    }
}

// one WARP-sized CTA per bit j (each on its own SM: the chains below are dependent point operations, and warps of one CTA
// would share a multiplier pipe): fold the slice sums of bit j (lanes), then 2^j by j doublings with the products of a
// doubling spread over the four lanes of a quad.  The last CTA to finish (ticket) adds the c values with a shuffle tree.
// (measured at n = 2^16, c = 13: 124 us as one 13-warp CTA -> 76 us; k_msm_bitsums 108 -> 52 us with 512-bucket slices)
__global__ void __launch_bounds__(32) k_msm_bitcombine(const xyzz_t* __restrict__ part, uint32_t bits, uint32_t slices,
                                                       xyzz_t* __restrict__ scaled /*[bits]*/, uint32_t* __restrict__ ticket,
                                                       xyzz_t* __restrict__ out) {
    const uint32_t bit = blockIdx.x, lane = threadIdx.x;
    xyzz_t v = xyzz_inf();
    for (uint32_t sidx = lane; sidx < slices; sidx += 32) v = xyzz_add_ni(v, part[bit * slices + sidx]);
    const uint32_t span = slices < 32 ? slices : 32;
#pragma unroll 1
    for (uint32_t off = 1; off < span; off <<= 1) v = xyzz_add_pair(v, (int)off);
    // every lane continues with lane 0's sum (the quad form needs whole warps; the quads compute the same value)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v.x.l[i] = __shfl_sync(0xffffffffu, v.x.l[i], 0);
        v.y.l[i] = __shfl_sync(0xffffffffu, v.y.l[i], 0);
        v.zz.l[i] = __shfl_sync(0xffffffffu, v.zz.l[i], 0);
        v.zzz.l[i] = __shfl_sync(0xffffffffu, v.zzz.l[i], 0);
    }
#pragma unroll 1
    for (uint32_t j = 0; j < bit; ++j) v = xyzz_dbl_quad(v);
    uint32_t last = 0;
    if (lane == 0) {
        scaled[bit] = v;
        __threadfence();
        last = atomicAdd(ticket, 1u) == bits - 1 ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    xyzz_t t = xyzz_inf();
    if (lane < bits) {
        const volatile uint32_t* src = reinterpret_cast<const volatile uint32_t*>(scaled + lane);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
        for (int i = 0; i < 32; ++i) dst[i] = src[i];
    }
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) t = xyzz_add_pair(t, off);
    if (lane == 0) out[0] = t;
}

// ---- two-level form for large bucket sets: with b = hi * l + lo (h groups of l buckets),
//          sum_b (b + 1) B_b = l * sum_hi hi S_hi + sum_lo (lo + 1) T_lo ,   S_hi = sum_lo B_(hi,lo) ,  T_lo = sum_hi B_(hi,lo)
//      — 2 nb additions like the running sums, but as h + l independent PLAIN sums (one CTA each: one addition per thread pair,
//      a pair-split shuffle tree, a shared-memory level), followed by the bit-parallel weighted sum above over only h + l
//      elements.  Depth ~10 + ~30 dependent operations instead of ~45 full additions on 64 warps, and the bulk of the
//      work runs at full occupancy.
static const int GS_THREADS = 64;   // (measured at 2^17 .. 2^20 points: 64 threads per group sum beat 32 and 128 by 0.3 - 1 %)
__global__ void __launch_bounds__(GS_THREADS) k_msm_group_sums(const xyzz_t* __restrict__ buckets, uint32_t h, uint32_t l,
                                                               xyzz_t* __restrict__ out /*[h + l]*/) {
    __shared__ xyzz_t sh[4];  // up to 128 threads
    const uint32_t g = blockIdx.x;
    const bool row = g < h;                       // S_g: l consecutive buckets; else T_(g - h): h buckets, stride l
    const uint32_t cnt = row ? l : h;
    const size_t base = row ? (size_t)g * l : (size_t)(g - h);
    const size_t stride = row ? 1 : l;
    xyzz_t acc = xyzz_inf();
#pragma unroll 1
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const xyzz_t* src = buckets + base + (size_t)i * stride;
        xyzz_t B;
        B.x = fp_load(&src->x);
        B.y = fp_load(&src->y);
        B.zz = fp_load(&src->zz);
        B.zzz = fp_load(&src->zzz);
        acc = xyzz_add_ni(acc, B);
    }
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) acc = xyzz_add_pair(acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        xyzz_t v = threadIdx.x < blockDim.x / 32 ? sh[threadIdx.x] : xyzz_inf();
#pragma unroll 1
        for (int off = 1; off < (int)(blockDim.x / 32); off <<= 1) v = xyzz_add_pair(v, off);
        if (threadIdx.x == 0) out[g] = v;
    }
}

// ---- running-sum form of the weighted bucket sum: 2 additions per bucket like the two-level form, but a longer chain on few
//      warps — the round-1 tail for large bucket sets, kept behind VKZG_MSM_TAIL=2 as the baseline of profiles/r02_msm_tail_sweep.txt
__device__ __noinline__ xyzz_t xyzz_mul_small(const xyzz_t p, uint32_t k) {
    xyzz_t acc = xyzz_inf();
    if (k == 0) return acc;
    int top = 31 - __clz(k);
    for (int b = top; b >= 0; --b) {
        acc = xyzz_dbl_ni(acc);
        if ((k >> b) & 1) acc = xyzz_add_ni(acc, p);
    }
    return acc;
}

static const int RED_SEG = 8;
static const int RED_THREADS = 128;

// sum_b (b + 1) B_b: thread = segment of RED_SEG buckets, CTA tree, one XYZZ per CTA
__global__ void __launch_bounds__(RED_THREADS) k_msm_reduce(const xyzz_t* __restrict__ buckets, uint32_t nb, xyzz_t* __restrict__ out) {
    __shared__ xyzz_t sh[RED_THREADS];
    uint32_t seg = blockIdx.x * RED_THREADS + threadIdx.x;
    uint32_t lo = seg * RED_SEG;
    xyzz_t run = xyzz_inf(), sum = xyzz_inf();
    if (lo < nb) {
        uint32_t hi = lo + RED_SEG < nb ? lo + RED_SEG : nb;
#pragma unroll 1
        for (uint32_t b = hi; b-- > lo;) {
            xyzz_t B;
            B.x = fp_load(&buckets[b].x);
            B.y = fp_load(&buckets[b].y);
            B.zz = fp_load(&buckets[b].zz);
            B.zzz = fp_load(&buckets[b].zzz);
            run = xyzz_add_ni(run, B);
            sum = xyzz_add_ni(sum, run);
        }
        // weights inside the segment were 1..RED_SEG; the true ones are lo + 1 .. lo + RED_SEG
        if (lo) sum = xyzz_add_ni(sum, xyzz_mul_small(run, lo));
    }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] = xyzz_add_ni(sh[threadIdx.x], sh[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) k_xyzz_sum(const xyzz_t* __restrict__ pts, uint32_t n, xyzz_t* __restrict__ out) {
    __shared__ xyzz_t sh[256];
    xyzz_t acc = xyzz_inf();
    for (uint32_t i = threadIdx.x; i < n; i += 256) acc = xyzz_add_ni(acc, pts[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] = xyzz_add_ni(sh[threadIdx.x], sh[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// h_scalars != nullptr: d_scalars is an allocated but EMPTY device buffer and the scalars are still in host memory — they are
// uploaded in pieces on the copy stream and every piece is scattered as soon as it has arrived (the scatter of piece k runs
// under the upload of piece k + 1; the cursors and lists are sized for the whole MSM, so the pieces simply append).
int32_t msm_large(vkzg_ctx* ctx, const Key& k, uint64_t first, const fp_t* d_scalars, uint64_t n, affine_t* d_out, const fp_t* h_scalars) {
    const uint32_t nb = 1u << (k.c - 1);
    DevBuf<uint32_t> counts, offsets, entries, order;
    DevBuf<xyzz_t> buckets, partial;
    cudaStream_t s = ctx->stream;
    VK_TRY(counts.alloc(ctx, nb + 2));  // [nb] = dropped entries of the optimistic scatter, [nb + 1] = bit-combine ticket
    VK_TRY(buckets.alloc(ctx, nb));
    VK_CUDA(cudaMemsetAsync(counts, 0, (nb + 2) * sizeof(uint32_t), s));
    uint32_t gb = ceil_div_u64(n, 256);
    // lanes per bucket: aim at ~32 additions per lane
    uint64_t avg = (uint64_t)n * k.W / nb;
    uint32_t P = 1;
    static int p_env = -1;
    if (p_env < 0) {
        const char* e = getenv("VKZG_MSM_P");
        p_env = e ? atoi(e) : 0;
    }
    while (P < 32 && avg / (P * 2) >= 48) P *= 2;  // 48 .. 95 additions per lane (P = 4 at n = 2^20, c = 17: 60 per lane; P = 2: -1 %, P = 8: -3 %)
    // (the top window is short — 7 bits at c = 13 — so the first ~100 buckets also receive n / 2^7 top digits each, 4x the mean
    //  list: with P = 32 their lanes still finish inside the kernel's throughput-bound time; striping those lists over extra
    //  lane groups was measured: it rescues P = 8 / 16 (1.03 -> 0.63 ms at 2^16) but the best configuration stays P = 32.)
    // small slices: one lane per bucket leaves most of the GPU idle (2^16 points: 32 K lanes of ~32 dependent additions on
    // 75 K thread slots) — split the buckets further until the grid fills ONE wave of lanes (two where the short top window
    // makes a few buckets several times longer than the rest: their lanes set the kernel's duration), down to ~6 additions per
    // lane.  Re-measured on the final kernels (profiles/r02_msm_lanes_sweep.txt): 2^17 points P = 8 (0.668 ms; 16: 0.730, 4: 0.721),
    // 2^18 P = 4 (1.016 ms; 8: 1.054, 2: 1.179), 2^19 P = 2 or 4, 2^16 (c = 13, hot top buckets) P = 32 (0.550 ms; 16: 0.691).
    const uint32_t top_shift0 = k.c * (k.W - 1);
    bool hot_top = false;
    if (top_shift0 < 254 && 254 - top_shift0 < k.c - 1) {
        const uint32_t r_top0 = top_shift0 >= 224 ? (0x30644e72u >> (top_shift0 - 224)) : 0xffffffffu;
        hot_top = n / (r_top0 ? r_top0 : 1) > avg / 2;
    }
    const uint64_t fill_lanes = (uint64_t)ctx->sm_count * 512 * (hot_top ? 2 : 1);
    while (P < 32 && (uint64_t)nb * P < fill_lanes && avg / (P * 2) >= 6) P *= 2;
    if (p_env == 1 || p_env == 2 || p_env == 4 || p_env == 8 || p_env == 16 || p_env == 32) P = (uint32_t)p_env;
    uint64_t threads = (uint64_t)nb * P;
    // ---- optimistic single pass (uniform scalars): fixed-capacity bucket lists, no count pass, no scan.  Entries that do
    //      not fit are counted in *dropped (device memory); the exact two-pass counting sort below is enqueued behind it and
    //      its kernels return at once unless *dropped != 0 (skewed scalars) — no host synchronisation either way.
    uint64_t cap64 = avg + avg / 4 + 64;
    // top window: 254 - c (W - 1) bits; its digits are <= (r >> shift) + 1 and uniform below that
    const uint32_t top_shift = k.c * (k.W - 1);
    uint32_t top_n = 0, top_extra = 0;
    if (top_shift < 254 && 254 - top_shift < k.c - 1) {
        const uint32_t r_top = top_shift >= 224 ? (0x30644e72u >> (top_shift - 224)) : 0xffffffffu;  // r >> top_shift (r < 2^254)
        top_n = r_top + 2 < nb ? r_top + 2 : nb;
        uint64_t per = n / (r_top ? r_top : 1);
        top_extra = (uint32_t)(per + per / 4 + 64);
    }
    // (A split form — scatter the first 1/k of the points, then scatter the rest on a high-priority side stream UNDER the bucket
    //  pass of the first part, second bucket pass on top of the first one's sums — was built and measured at 2^20 / 2^19 points,
    //  k = 2 .. 8: 3.00 - 3.26 ms against 2.90 ms unsplit (2^19: 1.77 - 2.08 against 1.63).  The two shorter bucket passes lose
    //  more to their partial last waves and to the scatter's CTAs than the 0.15 ms of hidden scatter returns.  Not kept.)
    const bool optimistic = n >= (1u << 12) && cap64 * nb + (uint64_t)top_n * top_extra < (1ull << 31) && !getenv("VKZG_MSM_TWO_PASS");
    const uint32_t* gate = nullptr;
    if (optimistic) {
        const uint32_t cap = (uint32_t)cap64;
        uint32_t* dropped = counts.p + nb;
        VK_TRY(entries.alloc(ctx, (size_t)cap * nb + (size_t)top_n * top_extra));
        if (h_scalars && n >= (1u << 17)) {
            ChunkedUpload up(ctx);
            VK_TRY(up.init());
            const uint64_t piece = ((n + 3) / 4 + 255) & ~255ull;
            for (uint64_t p0 = 0; p0 < n; p0 += piece) {
                const uint64_t np = n - p0 < piece ? n - p0 : piece;
                VK_TRY(up.copy(const_cast<fp_t*>(d_scalars) + p0, h_scalars + p0, np * sizeof(fp_t)));
                VK_TRY(up.publish());
                k_msm_scatter_fixed<<<ceil_div_u64(np, 256), 256, 0, s>>>(d_scalars + p0, np, k.c, k.W, k.n, first + p0, cap, top_n, top_extra,
                                                                        counts, entries, dropped);
                VK_TRY(launch_check(ctx));
            }
            h_scalars = nullptr;
        } else {
            if (h_scalars) {
                VK_CUDA(cudaMemcpyAsync(const_cast<fp_t*>(d_scalars), h_scalars, n * sizeof(fp_t), cudaMemcpyHostToDevice, s));
                h_scalars = nullptr;
            }
            k_msm_scatter_fixed<<<gb, 256, 0, s>>>(d_scalars, n, k.c, k.W, k.n, first, cap, top_n, top_extra, counts, entries, dropped);
            VK_TRY(launch_check(ctx));
        }
        const uint32_t* order_p = nullptr;
        if (!getenv("VKZG_MSM_NO_ORDER")) {  // (measurement knob)
            VK_TRY(order.alloc(ctx, nb));
            k_msm_order<<<1, 1024, 0, s>>>(counts, nb, order);
            VK_TRY(launch_check(ctx));
            order_p = order;
        }
        {
            KernelTimer timer(ctx);
            k_msm_bucket<<<ceil_div_u64(threads, 128), 128, 0, s>>>(k.table, counts, entries, nb, P, cap, top_n, top_extra, buckets, dropped, 0,
                                                                    order_p);
        }
        VK_TRY(launch_check(ctx));
        gate = dropped;
        k_msm_clear_if<<<ceil_div_u64(nb, 256), 256, 0, s>>>(counts, nb, gate);
        VK_TRY(launch_check(ctx));
    }
    // ---- exact two-pass counting sort (always enqueued; its kernels return at once behind a successful optimistic pass:
    //      ~10 us of empty launches instead of a device -> host -> device round trip).  Its n * W entry buffer (64 MB per
    //      2^20 points) comes from the context's pool like everything else.
    DevBuf<uint32_t> entries2;
    if (h_scalars && n) VK_CUDA(cudaMemcpyAsync(const_cast<fp_t*>(d_scalars), h_scalars, n * sizeof(fp_t), cudaMemcpyHostToDevice, s));
    VK_TRY(offsets.alloc(ctx, nb + 1));
    VK_TRY(entries2.alloc(ctx, (size_t)n * k.W));
    if (n) {
        k_msm_digits<false><<<gb, 256, 0, s>>>(d_scalars, n, k.c, k.W, k.n, first, counts, nullptr, nullptr, gate);
        VK_TRY(launch_check(ctx));
    }
    k_scan<<<1, 1024, 0, s>>>(counts, nb, offsets, gate);
    VK_TRY(launch_check(ctx));
    if (n) {
        k_msm_digits<true><<<gb, 256, 0, s>>>(d_scalars, n, k.c, k.W, k.n, first, counts, offsets, entries2, gate);
        VK_TRY(launch_check(ctx));
    }
    if (gate) {  // (the optimistic launch above is the timed one)
        k_msm_bucket<<<ceil_div_u64(threads, 128), 128, 0, s>>>(k.table, offsets, entries2, nb, P, 0, 0, 0, buckets, gate, 1, nullptr);
    } else {
        KernelTimer timer(ctx);
        k_msm_bucket<<<ceil_div_u64(threads, 128), 128, 0, s>>>(k.table, offsets, entries2, nb, P, 0, 0, 0, buckets, gate, 1, nullptr);
    }
    VK_TRY(launch_check(ctx));
    uint32_t rblocks;
    static int bitpar_max = -1, tail_form = -1;
    if (bitpar_max < 0) {
        const char* e = getenv("VKZG_MSM_BITPAR_MAX");
        bitpar_max = e ? atoi(e) : 4096;
        e = getenv("VKZG_MSM_TAIL");  // measurement knob: 1 = bit-parallel, 2 = running sums, 3 = two-level; default by size
        tail_form = e ? atoi(e) : 0;
    }
    int form = tail_form ? tail_form : (nb <= (uint32_t)bitpar_max ? 1 : 3);
    if (form == 3 && k.c < 5) form = 1;
    DevBuf<xyzz_t> groups;
    if (form == 1 || form == 3) {
        const xyzz_t* elems = buckets;
        uint32_t n_el = nb, bits = k.c, split_h = 0, split_l = 0;  // weights 1 .. 2^(c-1): c bits
        if (form == 3) {
            const uint32_t lg = k.c - 1;
            split_h = 1u << ((lg + 1) / 2);
            split_l = nb / split_h;
            n_el = split_h + split_l;
            VK_TRY(groups.alloc(ctx, n_el));
            static int gs_threads = -1;
            if (gs_threads < 0) {
                const char* e = getenv("VKZG_MSM_GS");  // measurement knob: threads per group sum (32 / 64 / 128)
                gs_threads = e ? atoi(e) : 0;
                if (gs_threads != 32 && gs_threads != 64 && gs_threads != 128) gs_threads = GS_THREADS;
            }
            k_msm_group_sums<<<n_el, gs_threads, 0, s>>>(buckets, split_h, split_l, groups);
            VK_TRY(launch_check(ctx));
            elems = groups;
            bits = lg;                                  // weights (h - 1) l < 2^(c-1) and l <= 2^((c-1)/2)
            while ((split_l >> bits) != 0) ++bits;      // (c - 1 = 1: l = 1 needs its own bit)
        }
        const uint32_t slices = (n_el + BS_SLICE - 1) / BS_SLICE;
        rblocks = bits * slices;
        VK_TRY(partial.alloc(ctx, (size_t)rblocks + 1 + bits));
        // (a thread per element for the single slice of the two-level form — 256 or 512 threads — was measured: 1 - 4 % slower,
        //  the wider shared-memory tree costs more than the <= 4 serial additions it removes)
        k_msm_bitsums<<<dim3(slices, bits), BS_THREADS, 0, s>>>(elems, n_el, slices, split_h, split_l, partial);
        VK_TRY(launch_check(ctx));
        k_msm_bitcombine<<<bits, 32, 0, s>>>(partial, bits, slices, partial.p + rblocks + 1, counts.p + nb + 1, partial.p + rblocks);
        VK_TRY(launch_check(ctx));
    } else {
        uint32_t segs = (nb + RED_SEG - 1) / RED_SEG;
        rblocks = (segs + RED_THREADS - 1) / RED_THREADS;
        VK_TRY(partial.alloc(ctx, rblocks + 1));
        k_msm_reduce<<<rblocks, RED_THREADS, 0, s>>>(buckets, nb, partial);
        VK_TRY(launch_check(ctx));
        k_xyzz_sum<<<1, 256, 0, s>>>(partial, rblocks, partial.p + rblocks);
        VK_TRY(launch_check(ctx));
    }
    return normalize_points(ctx, partial.p + rblocks, 1, d_out);
}

}  // namespace vk
