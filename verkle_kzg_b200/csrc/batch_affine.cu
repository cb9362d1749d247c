// Batch-affine summation for big batches of dense fixed-base MSMs (commits, IPA cross terms) — an OPT-IN path
// (VKZG_OPT_BATCH_AFFINE = 1), kept because it is correct, tested and the measured answer to "cut multiplications, not
// stalls" (VERDICT r1 item 3); the default stays the XYZZ kernel of commit.cu, which is faster on B200:
//
//     width-256 commits, batch 2^14, c = 20 (54.5 M table additions)      XYZZ 9.15 ms      batch-affine 10.35 ms
//     streaming levels 7.0 M additions/ms, gather level 4.6 M/ms, XYZZ kernel 5.96 M/ms   
//
// Why it was tried: the multiplier — not occupancy, not memory — bounds these kernels (profiles/r02_mulbench4_occupancy.txt:
// the 8 x 32-bit Montgomery product saturates at 67 G mul/s from two resident warps per sub-partition on), and an affine
// addition costs 6 field products (3 of them Montgomery's trick) against the 10 of the XYZZ mixed addition.
// Why it does not pay here: (1) the ~10 k ALU instructions of a shared inversion and the ~300 of the lazy subtractions,
// comparisons and index arithmetic per addition are NOT hidden under the multiplier — on this SM an ALU instruction
// issued next to IMAD.WIDE costs ~0.85 cycles (profiles/r02_ubench2_instruction_mix.txt), so a streamed addition costs
// ~5.3 k cycles per warp against 6.25 k in XYZZ, not 6/10 of it; (2) the first level has to gather its operands from the
// 100 GB window table for BOTH passes of the trick (or park them in local memory: +256 streamed bytes per pair), and it is
// half of all additions; (3) the tree adds an entry-list kernel, 6-7 level launches and ~400 streamed bytes per addition.
//
// An affine addition needs 1 / (x2 - x1); inversions only pay when thousands of INDEPENDENT additions share one.  Inside
// one job the additions into an accumulator are dependent — but the SUM of a job's table points is a tree: its n points
// are added pairwise (n/2 independent additions), the n/2 results pairwise again, ...  With thousands of jobs per call
// every level offers millions of independent additions, so a level is one kernel in which a thread chains K pairs through
// Montgomery's trick and the 32 lanes of a warp share ONE inversion (32 K additions per inversion).  Levels run until
// <= BA_TAIL points per job remain; the last few are summed in XYZZ by lanes + a shuffle tree (k_ba_tail).
//
//   k_ba_entries     scalars -> signed-digit table references, [job][w][term] (what commit.cu keeps in shared memory)
//   k_ba_level<1,K>  pairs of table points -> affine sums   (gathers once, parks the operands in local memory for pass 2)
//   k_ba_stream<K>   pairs of affine sums  -> affine sums   (64 registers: 8 warps per sub-partition cover the inversions)
//   k_ba_tail        <= BA_TAIL affine points per job -> XYZZ sum    (same output format as k_fixed_base_msm)
//
// Exceptional pairs never break a batch: an identity operand (zero digit, padding) passes the other point through, equal
// points take the tangent slope (denominator 2 y), opposite points give the identity — each of them with a denominator
// that is non-zero, so the shared product stays invertible.  Results are canonical affine points, hence bit-identical to
// the XYZZ path whatever the order of summation (tests/test_gpu_batch_affine.py compares both paths and the oracle).
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

static const uint32_t BA_NULL = 0xffffffffu;  // entry of a zero digit
static const uint32_t BA_TAIL = 32;           // points per job left to the XYZZ tail
static const int BA_K = 16;                   // pairs per thread and batch
static const int BA_THREADS = 128;

__device__ __forceinline__ affine_t ba_load_ro(const affine_t* p) {
    affine_t a;
    a.x = fp_load_ro(&p->x);
    a.y = fp_load_ro(&p->y);
    return a;
}

// one thread per (job, term): signed c-bit digits of the scalar -> W entries at [job][w][term]
__global__ void __launch_bounds__(256) k_ba_entries(const fp_t* __restrict__ scalars, uint32_t T, uint64_t jobs, uint32_t c, uint32_t W,
                                                    uint32_t ipa_m, uint32_t q_row, uint32_t* __restrict__ entries) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= jobs * T) return;
    const uint64_t job = i / T;
    const uint32_t term = (uint32_t)(i % T);
    fp_t k = fp_from_mont<S>(fp_load_ro(scalars + i));
    uint32_t base = term;
    if (ipa_m) {  // L / R cross terms of an IPA round (commit.cu: same base selection)
        const uint32_t side_off = (job & 1) ? 0u : ipa_m;
        base = (term == T - 1 && q_row != 0xffffffffu) ? q_row : (term / ipa_m) * 2 * ipa_m + side_off + term % ipa_m;
    }
    uint32_t* out = entries + job * ((uint64_t)T * W) + term;
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < W; ++w) {
        uint32_t v = scalar_bits(k.l, w * c, c) + carry;
        uint32_t neg = v >= half && w + 1 < W ? 1u : 0u;
        uint32_t mag = neg ? (1u << c) - v : v;
        carry = neg;
        out[(uint64_t)w * T] = mag ? ((((base * W + w) << (c - 1)) + (mag - 1)) | (neg << 31)) : BA_NULL;
    }
}

struct BaPair {
    affine_t a, b;
    int kind;  // 0 generic, 1 pass a, 2 pass b, 3 double a, 4 identity
};

template <int LEVEL0>
__device__ __forceinline__ affine_t ba_fetch(const affine_t* __restrict__ table, const uint32_t* __restrict__ entries,
                                             const affine_t* __restrict__ in, uint64_t idx) {
    if (LEVEL0) {
        uint32_t e = __ldg(entries + idx);
        if (e == BA_NULL) return affine_inf();
        affine_t p = ba_load_ro(table + (e & 0x7fffffffu));
        if ((e >> 31) && !affine_is_inf(p)) p.y = fp_neg<Q>(p.y);
        return p;
    }
    return ba_load_ro(in + idx);
}

__device__ __forceinline__ bool limbs_eq(const fp_t& a, const fp_t& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ b.l[i];
    return d == 0;
}

// the denominator of pair (a, b) and its kind; operands are canonical, so equality is limb equality
__device__ __forceinline__ fp_t ba_denominator(const affine_t& a, const affine_t& b, int& kind) {
    const bool ai = affine_is_inf(a), bi = affine_is_inf(b);
    if (ai || bi) {
        kind = ai ? (bi ? 4 : 2) : 1;
        return fp_one<Q>();
    }
    if (limbs_eq(a.x, b.x)) {
        if (limbs_eq(a.y, b.y) && !fp_is_zero(a.y)) {
            kind = 3;
            return fp_add_lazy<Q>(a.y, a.y);  // tangent: lambda = 3 x^2 / (2 y)
        }
        kind = 4;  // opposite points (or a 2-torsion point, which this curve does not have)
        return fp_one<Q>();
    }
    kind = 0;
    return fp_sub_lazy<Q>(b.x, a.x);
}

// One level of the pairwise tree.  n_in points per job in, n_out = ceil(n_in / 2) out; pair p of job j adds points 2p and
// 2p + 1 (the latter missing for odd n_in: the point passes through).  Thread t of a block owns the pairs
// base + k * BA_THREADS + t, k < K: neighbouring threads touch neighbouring points (coalesced 128-byte reads).
// The operands of pair k + 1 are in flight while pair k is processed (both passes).  LEVEL0 gathers its operands from the
// window table ONCE: pass 1 parks them in thread-local memory next to the running products (an interleaved, coalesced
// layout) and pass 2 reads them back — a second random walk over a 100 GB table costs far more than 128 streamed bytes.
struct BaCursor {
    uint64_t job;
    uint32_t p;
};
__device__ __forceinline__ void ba_advance(BaCursor& c, uint32_t step, uint32_t n_out) {
    c.p += step;
    while (c.p >= n_out) {
        c.p -= n_out;
        ++c.job;
    }
}
__device__ __forceinline__ void ba_retreat(BaCursor& c, uint32_t step, uint32_t n_out) {
    uint32_t s = step;
    while (s > c.p) {
        s -= c.p + 1;
        c.p = n_out - 1;
        --c.job;
    }
    c.p -= s;
}
template <int LEVEL0>
__device__ __forceinline__ void ba_fetch_pair(const affine_t* __restrict__ table, const uint32_t* __restrict__ entries,
                                              const affine_t* __restrict__ in, const BaCursor& c, uint32_t n_in, affine_t& a, affine_t& b) {
    const uint64_t i0 = c.job * n_in + 2 * (uint64_t)c.p;
    a = ba_fetch<LEVEL0>(table, entries, in, i0);
    b = 2 * c.p + 1 < n_in ? ba_fetch<LEVEL0>(table, entries, in, i0 + 1) : affine_inf();
}

template <int LEVEL0, int K>
__global__ void __launch_bounds__(BA_THREADS, 4) k_ba_level(const affine_t* __restrict__ table, const uint32_t* __restrict__ entries,
                                                            const affine_t* __restrict__ in, uint32_t n_in, uint32_t n_out,
                                                            uint64_t total_pairs, affine_t* __restrict__ out) {
    const int kmax = K;
    const uint64_t base = (uint64_t)blockIdx.x * (BA_THREADS * K) + threadIdx.x;
    fp_t pre[K];
    affine_t keep[LEVEL0 ? 2 * K : 1];
    fp_t run = fp_one<Q>();
    BaCursor cur;
    cur.job = base / n_out;
    cur.p = (uint32_t)(base % n_out);
    const BaCursor first = cur;
    const int nk = base >= total_pairs ? 0 : (int)((total_pairs - base + BA_THREADS - 1) / BA_THREADS < (uint64_t)kmax
                                                        ? (total_pairs - base + BA_THREADS - 1) / BA_THREADS
                                                        : (uint64_t)kmax);
    // ---- pass 1: running product of the denominators
    affine_t a, b, na, nb_;
    if (nk > 0) ba_fetch_pair<LEVEL0>(table, entries, in, cur, n_in, a, b);
#pragma unroll 1
    for (int k = 0; k < nk; ++k) {
        BaCursor nxt = cur;
        ba_advance(nxt, BA_THREADS, n_out);
        if (k + 1 < nk) ba_fetch_pair<LEVEL0>(table, entries, in, nxt, n_in, na, nb_);
        pre[k] = run;
        if (LEVEL0) {
            keep[2 * k] = a;
            keep[2 * k + 1] = b;
        }
        int kind;
        fp_t d = ba_denominator(a, b, kind);
        run = fp_mul_lazy_ni<Q>(run, d);
        a = na;
        b = nb_;
        cur = nxt;
    }
    // ---- one inversion per warp
    fp_t inv = warp_inverse_of_lane_products_t<Q>(fp_canon<Q>(run));
    // ---- pass 2: unwind the products, finish the additions (last pair first)
    cur = first;
    if (nk > 1) ba_advance(cur, (uint32_t)(nk - 1) * BA_THREADS % n_out, n_out), cur.job += (uint64_t)(nk - 1) * BA_THREADS / n_out;
    if (nk > 0) {
        if (LEVEL0) {
            a = keep[2 * (nk - 1)];
            b = keep[2 * (nk - 1) + 1];
        } else {
            ba_fetch_pair<0>(table, entries, in, cur, n_in, a, b);
        }
    }
#pragma unroll 1
    for (int k = nk - 1; k >= 0; --k) {
        BaCursor prv = cur;
        if (k > 0) {
            ba_retreat(prv, BA_THREADS, n_out);
            if (LEVEL0) {
                na = keep[2 * (k - 1)];
                nb_ = keep[2 * (k - 1) + 1];
            } else {
                ba_fetch_pair<0>(table, entries, in, prv, n_in, na, nb_);
            }
        }
        const uint64_t g = base + (uint64_t)k * BA_THREADS;
        int kind;
        fp_t d = ba_denominator(a, b, kind);
        fp_t dinv = fp_mul_lazy_ni<Q>(inv, pre[k]);
        inv = fp_mul_lazy_ni<Q>(inv, d);
        affine_t r;
        if (kind == 0 || kind == 3) {
            fp_t num;
            if (kind == 0) {
                num = fp_sub_lazy<Q>(b.y, a.y);
            } else {
                fp_t xx = fp_mul_lazy_ni<Q>(a.x, a.x);
                num = fp_add_lazy<Q>(fp_add_lazy<Q>(xx, xx), xx);
            }
            fp_t lam = fp_mul_lazy_ni<Q>(num, dinv);
            fp_t x3 = fp_sub_lazy<Q>(fp_sub_lazy<Q>(fp_mul_lazy_ni<Q>(lam, lam), a.x), b.x);
            fp_t y3 = fp_sub_lazy<Q>(fp_mul_lazy_ni<Q>(lam, fp_sub_lazy<Q>(a.x, x3)), a.y);
            r.x = fp_canon<Q>(x3);
            r.y = fp_canon<Q>(y3);
        } else if (kind == 1) {
            r = a;
        } else if (kind == 2) {
            r = b;
        } else {
            r = affine_inf();
        }
        fp_store(&out[g].x, r.x);
        fp_store(&out[g].y, r.y);
        a = na;
        b = nb_;
        cur = prv;
    }
}

// Streaming levels, register-lean form (the multiplier needs resident warps to cover the inversion phases: at 128
// registers — 4 warps per sub-partition — a third of them sit in the multiplier-free inversion at any time and the level
// ran at 53 % of the multiplier's rate).  No point of this curve has x = 0 (3 is a quadratic non-residue mod p), so an
// operand is the identity iff its x is 0 and pass 1 touches x coordinates only; the rare kinds (identity operand, equal x)
// take an out-of-line path.
__device__ __forceinline__ bool fp_limbs_zero(const fp_t& a) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i];
    return d == 0;
}
// denominator from the x coordinates alone; returns false for the rare kinds (y is needed)
__device__ __forceinline__ bool ba_den_x(const fp_t& ax, const fp_t& bx, fp_t& d) {
    if (fp_limbs_zero(ax) || fp_limbs_zero(bx) || limbs_eq(ax, bx)) return false;
    d = fp_sub_lazy<Q>(bx, ax);
    return true;
}
static __device__ __noinline__ fp_t ba_den_rare(const affine_t* __restrict__ in, uint64_t i0, bool has_b) {
    affine_t a = ba_load_ro(in + i0), b = has_b ? ba_load_ro(in + i0 + 1) : affine_inf();
    int kind;
    return ba_denominator(a, b, kind);
}
static __device__ __noinline__ void ba_add_rare(const affine_t* __restrict__ in, uint64_t i0, bool has_b, const fp_t dinv, affine_t* __restrict__ out) {
    affine_t a = ba_load_ro(in + i0), b = has_b ? ba_load_ro(in + i0 + 1) : affine_inf();
    int kind;
    ba_denominator(a, b, kind);
    affine_t r = affine_inf();
    if (kind == 1) r = a;
    if (kind == 2) r = b;
    if (kind == 3) {
        fp_t xx = fp_mul_lazy_ni<Q>(a.x, a.x);
        fp_t lam = fp_mul_lazy_ni<Q>(fp_add_lazy<Q>(fp_add_lazy<Q>(xx, xx), xx), dinv);
        fp_t x3 = fp_sub_lazy<Q>(fp_sub_lazy<Q>(fp_mul_lazy_ni<Q>(lam, lam), a.x), a.x);
        r.y = fp_canon<Q>(fp_sub_lazy<Q>(fp_mul_lazy_ni<Q>(lam, fp_sub_lazy<Q>(a.x, x3)), a.y));
        r.x = fp_canon<Q>(x3);
    }
    fp_store(&out->x, r.x);
    fp_store(&out->y, r.y);
}

template <int K, int MINB>
__global__ void __launch_bounds__(BA_THREADS, MINB) k_ba_stream(const affine_t* __restrict__ in, uint32_t n_in, uint32_t n_out,
                                                                uint64_t total_pairs, affine_t* __restrict__ out) {
    const uint64_t base = (uint64_t)blockIdx.x * (BA_THREADS * K) + threadIdx.x;
    fp_t pre[K];
    fp_t run = fp_one<Q>();
    BaCursor cur;
    cur.job = base / n_out;
    cur.p = (uint32_t)(base % n_out);
    const uint64_t left = base >= total_pairs ? 0 : (total_pairs - base + BA_THREADS - 1) / BA_THREADS;
    const int nk = left < (uint64_t)K ? (int)left : K;
#pragma unroll 1
    for (int k = 0; k < nk; ++k) {
        const uint64_t i0 = cur.job * n_in + 2 * (uint64_t)cur.p;
        const bool has_b = 2 * cur.p + 1 < n_in;
        pre[k] = run;
        fp_t d;
        bool plain = has_b;
        if (plain) {
            fp_t ax = fp_load_ro(&in[i0].x), bx = fp_load_ro(&in[i0 + 1].x);
            plain = ba_den_x(ax, bx, d);
        }
        if (!plain) d = ba_den_rare(in, i0, has_b);
        run = fp_mul_lazy_ni<Q>(run, d);
        ba_advance(cur, BA_THREADS, n_out);
    }
    fp_t inv = warp_inverse_of_lane_products_t<Q>(fp_canon<Q>(run));
#pragma unroll 1
    for (int k = nk - 1; k >= 0; --k) {
        ba_retreat(cur, BA_THREADS, n_out);
        const uint64_t i0 = cur.job * n_in + 2 * (uint64_t)cur.p;
        const bool has_b = 2 * cur.p + 1 < n_in;
        const uint64_t g = base + (uint64_t)k * BA_THREADS;
        fp_t dinv = fp_mul_lazy_ni<Q>(inv, pre[k]);
        fp_t d;
        bool plain = has_b;
        fp_t ax, bx;
        if (plain) {
            ax = fp_load_ro(&in[i0].x);
            bx = fp_load_ro(&in[i0 + 1].x);
            plain = ba_den_x(ax, bx, d);
        }
        if (!plain) {
            d = ba_den_rare(in, i0, has_b);
            inv = fp_mul_lazy_ni<Q>(inv, d);
            ba_add_rare(in, i0, has_b, dinv, out + g);
            continue;
        }
        inv = fp_mul_lazy_ni<Q>(inv, d);
        fp_t ay = fp_load_ro(&in[i0].y);
        fp_t lam = fp_mul_lazy_ni<Q>(fp_sub_lazy<Q>(fp_load_ro(&in[i0 + 1].y), ay), dinv);
        fp_t x3 = fp_sub_lazy<Q>(fp_sub_lazy<Q>(fp_mul_lazy_ni<Q>(lam, lam), ax), bx);
        fp_t y3 = fp_sub_lazy<Q>(fp_mul_lazy_ni<Q>(lam, fp_sub_lazy<Q>(ax, x3)), ay);
        fp_store(&out[g].x, fp_canon<Q>(x3));
        fp_store(&out[g].y, fp_canon<Q>(y3));
    }
}

// the last <= BA_TAIL points of every job: LPJ lanes per job, XYZZ mixed additions + a pair-split shuffle tree
template <int LPJ>
__global__ void __launch_bounds__(128, 4) k_ba_tail(const affine_t* __restrict__ in, uint32_t n_in, uint64_t jobs, xyzz_t* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t job = t / LPJ;
    const uint32_t gl = (uint32_t)(t % LPJ);
    const bool live = job < jobs;
    xyzz_t acc = xyzz_inf();
    if (live) {
        for (uint32_t i = gl; i < n_in; i += LPJ) {
            affine_t p = ba_load_ro(in + job * n_in + i);
            xyzz_madd_hot(acc, p);
        }
    }
    xyzz_canon(acc);
#pragma unroll 1
    for (int off = LPJ / 2; off > 0; off >>= 1) acc = xyzz_add_pair(acc, off);
    if (gl == 0 && live) {
        fp_store(&out[job].x, acc.x);
        fp_store(&out[job].y, acc.y);
        fp_store(&out[job].zz, acc.zz);
        fp_store(&out[job].zzz, acc.zzz);
    }
}

// jobs x T-term dense fixed-base MSMs through the batch-affine tree (same contract as fixed_base_msm)
int32_t fixed_base_msm_batch_affine(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                                    uint32_t q_row, xyzz_t* d_out) {
    if (jobs == 0) return VKZG_OK;
    cudaStream_t s = ctx->stream;
    const uint64_t n0 = (uint64_t)T * k.W;
    if (n0 >= (1ull << 31) || jobs * n0 >= (1ull << 40)) return VKZG_ERR_RANGE;
    DevBuf<uint32_t> entries;
    DevBuf<affine_t> bufA, bufB;
    VK_TRY(entries.alloc(ctx, jobs * n0));
    const uint64_t n1 = (n0 + 1) / 2, n2 = (n1 + 1) / 2;
    VK_TRY(bufA.alloc(ctx, jobs * n1));
    VK_TRY(bufB.alloc(ctx, jobs * n2));
    KernelTimer timer(ctx);  // (the whole tree counts as the dominant kernel of the call)
    k_ba_entries<<<ceil_div_u64(jobs * T, 256), 256, 0, s>>>(d_scalars, T, jobs, k.c, k.W, ipa_m, q_row, entries);
    VK_TRY(launch_check(ctx));
    const uint32_t per_block = BA_THREADS * BA_K;
    static int k_env = -1;
    if (k_env < 0) {
        const char* e = getenv("VKZG_BA_K");  // tuning knob: pairs per thread in the streaming levels (16 or 32)
        k_env = e ? atoi(e) : 16;
    }
    uint32_t n_in = (uint32_t)n0;
    affine_t* src = nullptr;
    affine_t* dst = bufA;
    bool first = true;
    while (first || n_in > BA_TAIL) {
        const uint32_t n_out = (n_in + 1) / 2;
        const uint64_t total = jobs * n_out;
        const uint32_t pb = (!first && k_env == 32) ? BA_THREADS * 32 : per_block;
        const uint64_t blocks = (total + pb - 1) / pb;
        if (blocks > 0x7fffffffull) return VKZG_ERR_RANGE;
        if (first)
            k_ba_level<1, BA_K><<<(uint32_t)blocks, BA_THREADS, 0, s>>>(k.table, entries, nullptr, n_in, n_out, total, dst);
        else if (k_env == 32)
            k_ba_stream<32, 8><<<(uint32_t)blocks, BA_THREADS, 0, s>>>(src, n_in, n_out, total, dst);
        else
            k_ba_stream<16, 8><<<(uint32_t)blocks, BA_THREADS, 0, s>>>(src, n_in, n_out, total, dst);
        VK_TRY(launch_check(ctx));
        first = false;
        src = dst;
        dst = dst == bufA.p ? bufB.p : bufA.p;
        n_in = n_out;
    }
    k_ba_tail<8><<<ceil_div_u64(jobs * 8, 128), 128, 0, s>>>(src, n_in, jobs, d_out);
    return launch_check(ctx);
}

}  // namespace vk
