// M1 for width-N vectors: `jobs` independent fixed-base MSMs, one warp each.
//
// utils::inner_product (utils.rs:16-19) multiplies every base by its scalar with a 254-bit
// double-and-add.  The bases of a key never change, so here every term is W table look-ups instead:
// the scalar is recoded into W signed c-bit digits d_w and  s * P = sum_w T[P][w][|d_w|] * sign(d_w).
// A warp turns a chunk of 128 scalars into a compact list of (table index, sign) entries in shared
// memory, its lanes then walk that list 32 entries at a time — perfectly balanced whatever the number
// of terms or zero digits — each lane adding into its own XYZZ accumulator with a 64-byte vectorised
// gather per addition (the next entry's point is in flight while the current one is added).  The 32
// accumulators are summed with a shuffle tree.
#include "vk_common.cuh"
#include "warp_util.cuh"

namespace vk {

__device__ __forceinline__ affine_t load_affine_ro(const affine_t* p) {
    affine_t a;
    a.x = fp_load_ro(&p->x);
    a.y = fp_load_ro(&p->y);
    return a;
}

__device__ __forceinline__ xyzz_t shfl_down_xyzz(const xyzz_t& v, int off) {
    xyzz_t r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], off);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], off);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], off);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], off);
    }
    return r;
}

// Recode one canonical scalar into signed c-bit digits and append the non-zero ones to `list`.
__device__ __forceinline__ void emit_entries(const fp_t& k, uint32_t row_base /* base * W */, uint32_t c, uint32_t W,
                                             uint32_t* list, uint32_t* cnt) {
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < W; ++w) {
        uint32_t v = scalar_bits(k.l, w * c, c) + carry;
        uint32_t neg = v >= half && w + 1 < W ? 1u : 0u;  // the top window never needs to borrow (k < r < 2^254)
        uint32_t mag = neg ? (1u << c) - v : v;
        carry = neg;
        if (mag) {
            uint32_t pos = atomicAdd(cnt, 1u);
            list[pos] = (((row_base + w) << (c - 1)) + (mag - 1)) | (neg << 31);
        }
    }
}

// LPJ lanes per job (32, 16, 8, 4, 2 or 1): a warp runs 32 / LPJ jobs side by side, each group of LPJ lanes with its own
// slice of the shared-memory entry list.  Wide jobs (commits, IPA cross terms) use the whole warp; verkle nodes
// with a handful of terms use 4 or 8 lanes so that the shuffle-tree fold (log2 LPJ full additions) does not
// dominate their few table additions.
template <int LPJ>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
    k_fixed_base_msm(const affine_t* __restrict__ table, uint32_t c, uint32_t W, const fp_t* __restrict__ scalars, uint32_t T,
                     uint64_t jobs, uint32_t ipa_m, uint32_t q_row, const uint32_t* __restrict__ row_ptr,
                     const uint16_t* __restrict__ slot, uint32_t split, xyzz_t* __restrict__ out) {
    extern __shared__ uint32_t smem[];
    constexpr uint32_t JPW = 32 / LPJ;                 // jobs per warp
    constexpr uint32_t CHUNK = CHUNK_TERMS / JPW;      // terms recoded per pass and group
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gl = lane % LPJ, grp = lane / LPJ;
    const uint32_t list_cap = CHUNK * W;
    uint32_t* list = smem + warp * (CHUNK_TERMS * W + 32) + grp * list_cap;
    uint32_t* cnt = smem + warp * (CHUNK_TERMS * W + 32) + CHUNK_TERMS * W + grp;
    // split > 1 (few wide jobs, e.g. a single proof): `jobs` counts SLICES, slice v covers the terms
    // [s * Tsub, (s + 1) * Tsub) of job v / split and writes a partial sum (k_sum_slices adds them up)
    const uint64_t vjob = ((uint64_t)blockIdx.x * WARPS_PER_CTA + warp) * JPW + grp;
    const bool live = vjob < jobs;
    uint64_t job = vjob;
    uint32_t term0 = 0;
    if (split > 1) job = vjob / split;
    // dense: job j owns scalars[j*T .. (j+1)*T), term t uses base t.  CSR (row_ptr != nullptr, verkle nodes):
    // job j owns terms [row_ptr[j], row_ptr[j+1]) and term t uses base slot[t].
    const fp_t* sc = scalars + job * T;
    const uint16_t* sl = nullptr;
    if (row_ptr && live) {
        uint32_t t0 = row_ptr[job];
        T = row_ptr[job + 1] - t0;
        sc = scalars + t0;
        sl = slot + t0;
    }
    const uint32_t Tfull = T;
    if (split > 1) {
        uint32_t Tsub = (T + split - 1) / split;
        term0 = (uint32_t)(vjob % split) * Tsub;
        T = term0 >= T ? 0 : (T - term0 < Tsub ? T - term0 : Tsub);  // terms of this slice
        sc += term0;
        if (sl) sl += term0;
    }
    if (!live) T = 0;
    // ipa_m != 0: job 2p is the L cross term of proof p, job 2p+1 the R cross term of an IPA round whose
    // half length is ipa_m (see ipa.cu): term j < T-1 uses base (j / m) * 2m + (L ? m : 0) + j % m and the
    // last term the base q_row (when q_row != 0xffffffff)
    const uint32_t side_off = (job & 1) ? 0u : ipa_m;
    // every lane of the warp runs the same number of passes (the __syncwarp()s below must be convergent)
    uint32_t Tmax = T;
    if (JPW > 1) {
#pragma unroll
        for (int m = 16; m >= LPJ; m >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, m));
    }

    xyzz_t acc = xyzz_inf();
    for (uint32_t chunk = 0; chunk < Tmax; chunk += CHUNK) {
        if (gl == 0) *cnt = 0;
        __syncwarp();
        for (uint32_t j = gl; j < CHUNK; j += LPJ) {
            uint32_t term = chunk + j;
            if (term < T) {
                fp_t k = fp_from_mont<S>(fp_load_ro(sc + term));
                uint32_t gt = term0 + term;  // term index within the whole job
                uint32_t base = sl ? (uint32_t)sl[term] : gt;
                if (ipa_m) base = (gt == Tfull - 1 && q_row != 0xffffffffu) ? q_row : (gt / ipa_m) * 2 * ipa_m + side_off + gt % ipa_m;
                emit_entries(k, base * W, c, W, list, cnt);
            }
        }
        __syncwarp();
        const uint32_t n = *cnt;
        uint32_t t = gl;
        uint32_t e = 0;
        affine_t cur;
        if (t < n) {
            e = list[t];
            cur = load_affine_ro(table + (e & 0x7fffffffu));
        }
        while (t < n) {
            uint32_t tn = t + LPJ, en = 0;
            affine_t nxt;
            if (tn < n) {
                en = list[tn];
                nxt = load_affine_ro(table + (en & 0x7fffffffu));
            }
            if (e >> 31) cur.y = fp_neg<Q>(cur.y);
            xyzz_madd_hot(acc, cur);
            cur = nxt;
            e = en;
            t = tn;
        }
        __syncwarp();
    }
    xyzz_canon(acc);
#pragma unroll 1
    for (int off = LPJ / 2; off > 0; off >>= 1) acc = xyzz_add_pair(acc, off);
    if (gl == 0 && live) {
        fp_store(&out[vjob].x, acc.x);
        fp_store(&out[vjob].y, acc.y);
        fp_store(&out[vjob].zz, acc.zz);
        fp_store(&out[vjob].zzz, acc.zzz);
    }
}

// out[j] = sum of the `split` slice sums of job j (one warp per job, shuffle tree)
__global__ void __launch_bounds__(128) k_sum_slices(const xyzz_t* __restrict__ part, uint64_t jobs, uint32_t split, xyzz_t* __restrict__ out) {
    uint64_t job = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (job >= jobs) return;
    xyzz_t acc = xyzz_inf();
    for (uint32_t s = lane; s < split; s += 32) acc = xyzz_add_ni(acc, part[job * split + s]);
    int top = 16;  // only as many levels as there are slices
    while (top > 1 && (uint32_t)top >= split) top >>= 1;
#pragma unroll 1
    for (int off = top; off > 0; off >>= 1) acc = xyzz_add_pair(acc, off);
    if (lane == 0) out[job] = acc;
}

template <int LPJ>
static int32_t launch_fixed_base(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                                 uint32_t q_row, const uint32_t* d_row_ptr, const uint16_t* d_slot, uint32_t split, xyzz_t* d_out) {
    static bool attr_set[64] = {false};  // function attributes are per device: one process may hold contexts on several GPUs
    if (ctx->device >= 64 || !attr_set[ctx->device]) {
        VK_CUDA(cudaFuncSetAttribute(k_fixed_base_msm<LPJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VK_CUDA(cudaFuncSetAttribute(k_fixed_base_msm<LPJ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        if (ctx->device < 64) attr_set[ctx->device] = true;
    }
    size_t smem = (size_t)WARPS_PER_CTA * (CHUNK_TERMS * k.W + 32) * sizeof(uint32_t);
    uint64_t per_cta = (uint64_t)WARPS_PER_CTA * (32 / LPJ);
    uint64_t blocks = (jobs + per_cta - 1) / per_cta;
    if (blocks > 0x7fffffffull) return VKZG_ERR_RANGE;
    KernelTimer timer(ctx);
    k_fixed_base_msm<LPJ><<<(uint32_t)blocks, WARPS_PER_CTA * 32, smem, ctx->stream>>>(k.table, k.c, k.W, d_scalars, T, jobs, ipa_m, q_row,
                                                                                       d_row_ptr, d_slot, split, d_out);
    return launch_check(ctx);
}

// lanes_per_job: 0 = whole warp.  csr_split > 1 (CSR only): every node's term list is sliced over that many warps.
int32_t fixed_base_msm_csr(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                           uint32_t q_row, const uint32_t* d_row_ptr, const uint16_t* d_slot, xyzz_t* d_out, uint32_t lanes_per_job,
                           uint32_t csr_split) {
    if (jobs == 0) return VKZG_OK;
    if (d_row_ptr && csr_split > 1) {
        DevBuf<xyzz_t> part;
        VK_TRY(part.alloc(ctx, jobs * csr_split));
        VK_TRY(launch_fixed_base<32>(ctx, k, d_scalars, T, jobs * csr_split, 0, 0xffffffffu, d_row_ptr, d_slot, csr_split, part));
        k_sum_slices<<<ceil_div_u64(jobs * 32, 128), 128, 0, ctx->stream>>>(part, jobs, csr_split, d_out);
        return launch_check(ctx);
    }
    if (lanes_per_job == 1) return launch_fixed_base<1>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, d_row_ptr, d_slot, 1, d_out);
    if (lanes_per_job == 2) return launch_fixed_base<2>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, d_row_ptr, d_slot, 1, d_out);
    if (lanes_per_job == 4) return launch_fixed_base<4>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, d_row_ptr, d_slot, 1, d_out);
    if (lanes_per_job == 8) return launch_fixed_base<8>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, d_row_ptr, d_slot, 1, d_out);
    // big dense batches of wide jobs: the batch-affine tree (batch_affine.cu)
    if (!d_row_ptr && lanes_per_job == 0) {
        static int ba_env = -2;
        if (ba_env == -2) {
            const char* e = getenv("VKZG_BATCH_AFFINE");
            ba_env = e ? atoi(e) : -1;
        }
        const int mode = ctx->batch_affine >= 0 ? ctx->batch_affine : ba_env;
        const bool big = jobs * (uint64_t)T * k.W >= (4ull << 20);  // >= 4 M additions: every level keeps thousands of warps busy
        // (an explicit VKZG_OPT_BATCH_AFFINE = 1 is the caller's decision whatever the size — the tests use small batches)
        if (mode == 1 && (big || ctx->batch_affine == 1)) return fixed_base_msm_batch_affine(ctx, k, d_scalars, T, jobs, ipa_m, q_row, d_out);
    }
    // big dense batches: fewer lanes per job shorten the shuffle-tree fold (log2(lanes) full additions per job) as long as
    // enough warps remain to fill the GPU several times over
    if (!d_row_ptr && lanes_per_job == 0) {
        static int dense_lpj = -1;
        if (dense_lpj < 0) {
            const char* e = getenv("VKZG_FB_LPJ");  // tuning knob (8, 16 or 32)
            dense_lpj = e ? atoi(e) : 8;
        }
        // minimum number of half-waves of warps: 20 for a lone launch (a grid of a few waves of long warps leaves the SMs idle
        // in its tail; with the pair-split fold the extra fold levels of whole-warp jobs cost less than that: commits at
        // 2^14 measured 1.790 M/s with 32 lanes per job, 1.772 M/s with 16 or 8), 3 for the IPA cross terms, whose
        // half-batches run on two streams and fill each other's tails (measured: 196.6 k proofs/s with 8 lanes, 188.2 k with 32)
        static int fill_env = -1;
        if (fill_env < 0) {
            const char* e = getenv("VKZG_FB_FILL_X2");
            fill_env = e ? atoi(e) : 0;
        }
        const int fill_x2 = fill_env ? fill_env : (ipa_m && ctx->ipa_two_streams ? 3 : 20);
        uint64_t warps16 = jobs / 2, warps8 = jobs / 4, fill = (uint64_t)ctx->sm_count * 16 * fill_x2 / 6;
        if (dense_lpj == 8 && warps8 >= 3 * fill) return launch_fixed_base<8>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, nullptr, nullptr, 1, d_out);
        if (dense_lpj <= 16 && warps16 >= 3 * fill) return launch_fixed_base<16>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, nullptr, nullptr, 1, d_out);
    }
    // few wide dense jobs (single proofs / commits): slice every job over several warps so the whole GPU works on it
    if (!d_row_ptr && T >= 16 && jobs * 8 <= (uint64_t)ctx->sm_count * 16) {
        // latency-bound: ~4 terms (2-4 additions per lane) per warp, at most 32 slices so that k_sum_slices is one shuffle
        // tree — every halving of the per-lane chain costs one more 7-product fold level, which is where it stops paying
        uint32_t split = (T + 3) / 4;
        uint64_t room = (uint64_t)ctx->sm_count * 16 / jobs;
        if (split > room) split = (uint32_t)room;
        if (split > 32) split = 32;
        if (split > 1) {
            DevBuf<xyzz_t> part;
            VK_TRY(part.alloc(ctx, jobs * split));
            VK_TRY(launch_fixed_base<32>(ctx, k, d_scalars, T, jobs * split, ipa_m, q_row, nullptr, nullptr, split, part));
            k_sum_slices<<<ceil_div_u64(jobs * 32, 128), 128, 0, ctx->stream>>>(part, jobs, split, d_out);
            return launch_check(ctx);
        }
    }
    return launch_fixed_base<32>(ctx, k, d_scalars, T, jobs, ipa_m, q_row, d_row_ptr, d_slot, 1, d_out);
}

int32_t fixed_base_msm(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                       uint32_t q_row, xyzz_t* d_out) {
    return fixed_base_msm_csr(ctx, k, d_scalars, T, jobs, ipa_m, q_row, nullptr, nullptr, d_out, 0, 1);
}

// -------------------------------------------------------------------------------------------------
// sum of n affine points -> one affine point (combine step for sharded MSMs); single CTA tree.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_g1_sum(const affine_t* __restrict__ pts, uint64_t n, xyzz_t* __restrict__ out) {
    __shared__ xyzz_t sh[256];
    xyzz_t acc = xyzz_inf();
    for (uint64_t i = threadIdx.x; i < n; i += 256) xyzz_madd(acc, pts[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] = xyzz_add_ni(sh[threadIdx.x], sh[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

int32_t g1_sum(vkzg_ctx* ctx, const affine_t* d_points, uint64_t n, affine_t* d_out) {
    DevBuf<xyzz_t> acc;
    VK_TRY(acc.alloc(ctx, 1));
    k_g1_sum<<<1, 256, 0, ctx->stream>>>(d_points, n, acc);
    VK_TRY(launch_check(ctx));
    return normalize_points(ctx, acc, 1, d_out);
}

// -------------------------------------------------------------------------------------------------
// D1: VCCommitment::to_data_item (vector-commit/src/lib.rs:56-67): identity -> 0, else the 32 compressed
// bytes (flag bits included, quirk Q7) reduced mod r.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_to_data_item(const affine_t* __restrict__ pts, uint64_t n, fp_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t p;
    p.x = fp_load(&pts[i].x);
    p.y = fp_load(&pts[i].y);
    fp_t r = fp_zero<S>();
    if (!affine_is_inf(p)) {
        fp_t v, r2;
        affine_compress(p, v.l);
#pragma unroll
        for (int k = 0; k < 8; ++k) r2.l[k] = S::r2(k);
        r = fp_mul<S>(r2, v);
    }
    fp_store(out + i, r);
}

int32_t to_data_item(vkzg_ctx* ctx, const affine_t* d_points, uint64_t n, fp_t* d_out) {
    if (n == 0) return VKZG_OK;
    k_to_data_item<<<ceil_div_u64(n, 128), 128, 0, ctx->stream>>>(d_points, n, d_out);
    return launch_check(ctx);
}

}  // namespace vk
