// C ABI of libvkzg.so (include/vkzg.h): context / key management, the M1 + D1 entry points and the
// measurement probes.  The scheme-level entry points live in ipa.cu, poly.cu, multiproof.cu, tree.cu.
#include "vk_common.cuh"
#include "field_kara.cuh"  // the probe of the Karatsuba product (measured, not used by the kernels)

using namespace vk;


extern "C" {

const char* vkzg_strerror(int32_t s) {
    switch (s) {
        case VKZG_OK: return "ok";
        case VKZG_ERR_CUDA: return "CUDA error or no sm_100 device (there is no CPU fallback)";
        case VKZG_ERR_ARG: return "invalid argument";
        case VKZG_ERR_RANGE: return "index or width outside the key";
        case VKZG_ERR_UNSUPPORTED: return "unsupported shape";
        case VKZG_ERR_OOM: return "out of device memory";
        default: return "unknown status";
    }
}

uint32_t vkzg_abi_version(void) { return 1; }

static int32_t ctx_create(vkzg_ctx** out, int32_t device_id, void* cuda_stream, bool own) {
    if (!out) return VKZG_ERR_ARG;
    *out = nullptr;
    int count = 0;
    VK_CUDA(cudaGetDeviceCount(&count));
    if (device_id < 0 || device_id >= count) return VKZG_ERR_ARG;
    VK_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    VK_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10) {
        fprintf(stderr, "[vkzg] device %d is sm_%d%d; this library is built for sm_100a only\n", device_id, prop.major, prop.minor);
        return VKZG_ERR_CUDA;
    }
    vkzg_ctx* ctx = new vkzg_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (!own) {
        ctx->stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return VKZG_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    // scratch memory comes from a pool PRIVATE to this context (the device's default pool and its release threshold are the
    // host application's — torch, another library — and are left alone).  Freed scratch stays cached in the pool between
    // calls up to VKZG_POOL_KEEP_MB (environment; default: everything) and goes back to the driver at vkzg_ctx_trim /
    // vkzg_ctx_destroy.
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device_id;
    if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess) {
        cudaGetLastError();
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return VKZG_ERR_CUDA;
    }
    uint64_t thr = ~0ull;
    if (const char* e = getenv("VKZG_POOL_KEEP_MB")) thr = (uint64_t)strtoull(e, nullptr, 10) << 20;
    cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    *out = ctx;
    return VKZG_OK;
}

int32_t vkzg_ctx_create(vkzg_ctx** out, int32_t device_id) { return ctx_create(out, device_id, nullptr, true); }
int32_t vkzg_ctx_create_on_stream(vkzg_ctx** out, int32_t device_id, void* cuda_stream) {
    return ctx_create(out, device_id, cuda_stream, false);
}

int32_t vkzg_ctx_destroy(vkzg_ctx* ctx) {
    if (!ctx) return VKZG_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->keys) {
        cudaFree(kv.second.bases);
        cudaFree(kv.second.table);
        cudaFree(kv.second.dom.omega);
    }
    for (auto& kv : ctx->domains) cudaFree(kv.second.omega);
    for (void* h : ctx->host_stage)
        if (h) cudaFreeHost(h);
    for (auto& ev : ctx->timing_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (auto st : ctx->side_streams) cudaStreamSynchronize(st);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);  // every scratch block goes back to the driver
    for (auto st : ctx->side_streams) cudaStreamDestroy(st);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VKZG_OK;
}

int32_t vkzg_ctx_sync(vkzg_ctx* ctx) {
    VK_TRY(ctx_check(ctx));
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    return VKZG_OK;
}

int32_t vkzg_ctx_trim(vkzg_ctx* ctx) {
    VK_TRY(ctx_check(ctx));
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->aux_stream) VK_CUDA(cudaStreamSynchronize(ctx->aux_stream));
    if (ctx->copy_stream) VK_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    for (auto st : ctx->side_streams) VK_CUDA(cudaStreamSynchronize(st));
    VK_CUDA(cudaMemPoolTrimTo(ctx->pool, 0));
    return VKZG_OK;
}

uint64_t vkzg_ctx_launches(const vkzg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t vkzg_ctx_set_option(vkzg_ctx* ctx, int32_t option, int32_t value) {
    VK_TRY(ctx_check(ctx));
    switch (option) {
        case VKZG_OPT_IPA_TWO_STREAMS: ctx->ipa_two_streams = value != 0; return VKZG_OK;
        case VKZG_OPT_TREE_FLATTEN:
            if (value < 0 || value > 2) return VKZG_ERR_ARG;
            ctx->tree_flatten = value;
            return VKZG_OK;
        case VKZG_OPT_MULTIPROOF_CHECK_Y: ctx->multiproof_check_y = value != 0; return VKZG_OK;
        case VKZG_OPT_BATCH_AFFINE:
            if (value < -1 || value > 1) return VKZG_ERR_ARG;
            ctx->batch_affine = value;
            return VKZG_OK;
        default: return VKZG_ERR_ARG;
    }
}

int32_t vkzg_ctx_kernel_timing(vkzg_ctx* ctx, int32_t enable) {
    VK_TRY(ctx_check(ctx));
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& ev : ctx->timing_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    ctx->timing_events.clear();
    ctx->timing = enable != 0;
    return VKZG_OK;
}

int32_t vkzg_ctx_kernel_timing_read(vkzg_ctx* ctx, uint64_t* launches, double* total_ms) {
    VK_TRY(ctx_check(ctx));
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    double ms = 0;
    for (auto& ev : ctx->timing_events) {
        float t = 0;
        VK_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
        ms += t;
    }
    if (launches) *launches = ctx->timing_events.size();
    if (total_ms) *total_ms = ms;
    return VKZG_OK;
}

int32_t vkzg_key_load_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_bases, uint32_t n, const vkzg_g1_affine* d_q, uint32_t kind,
                          uint32_t window_bits, uint32_t* key_id) {
    VK_TRY(ctx_check(ctx));
    if (!d_bases || !n || !key_id) return VKZG_ERR_ARG;
    if (kind != VKZG_KEY_WINDOW && kind != VKZG_KEY_MSM) return VKZG_ERR_ARG;
    Key k;
    k.kind = kind;
    k.n = n;
    k.has_q = d_q != nullptr;
    if (kind == VKZG_KEY_MSM && d_q) return VKZG_ERR_ARG;
    uint32_t c = window_bits;
    // MSM keys: the weighted bucket sum is a latency-bound tail whose cost grows with the 2^(c-1) buckets, the bucket pass with
    // the ceil(255/c) digits per scalar.  Measured on B200 (profiles/r02_msm_window_sweep.txt, profiles/r02_msm_tail_sweep.txt):
    // with the two-level tail c = 17 (15 digits) from 2^19 points on, 16 at 2^18, 15 at 2^17 — the slices of a sharded MSM —,
    // c = 13 below.  (A width whose TOP window is only a few bits wide — c = 14, 18, 19 — sends a large share of all scalars
    // into a handful of buckets: correct but 20-250x slower; never the default.)
    if (c == 0)
        c = kind == VKZG_KEY_WINDOW ? 16
            : (n >= (1u << 19) ? 17 : (n >= (1u << 18) ? 16 : (n >= (1u << 17) ? 15 : (n >= (1u << 14) ? 13 : 12))));
    if (kind == VKZG_KEY_MSM && window_bits == 0 && getenv("VKZG_MSM_C")) c = (uint32_t)atoi(getenv("VKZG_MSM_C"));  // measurement knob
    if (c < 2 || c > 20) return VKZG_ERR_ARG;
    k.c = c;
    k.W = (255 + c - 1) / c;  // scalars are < r < 2^254: 254 bits + the carry of the signed recoding
    // the fixed-base kernel keeps CHUNK_TERMS * W list entries per warp in shared memory: very narrow windows do not fit
    if (kind == VKZG_KEY_WINDOW && fixed_base_smem_bytes(k.W) > 200 * 1024) return VKZG_ERR_ARG;
    uint32_t nb = n + (k.has_q ? 1 : 0);
    if (kind == VKZG_KEY_WINDOW) {
        if (((uint64_t)nb * k.W) << (c - 1) >= (1ull << 31)) return VKZG_ERR_RANGE;
        if (nb > 65535) return VKZG_ERR_RANGE;
    } else {
        if ((uint64_t)n * k.W >= (1ull << 31)) return VKZG_ERR_RANGE;
    }
    VK_CUDA(cudaMalloc((void**)&k.bases, (size_t)nb * sizeof(affine_t)));
    int32_t st = VKZG_OK;
    if (cudaMemcpyAsync(k.bases, d_bases, (size_t)n * sizeof(affine_t), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess ||
        (d_q && cudaMemcpyAsync(k.bases + n, d_q, sizeof(affine_t), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)) {
        cudaGetLastError();
        st = VKZG_ERR_CUDA;
    }
    // every base must be the identity (0,0) or a canonical point of the curve: one off-curve base (a zero denominator in
    // the batched table additions) would otherwise corrupt the tables of unrelated rows through the shared inversions
    if (st == VKZG_OK) st = check_points_on_curve(ctx, k.bases, nb);
    if (st == VKZG_OK) st = kind == VKZG_KEY_WINDOW ? build_window_tables(ctx, k) : build_msm_tables(ctx, k);
    if (st == VKZG_OK && kind == VKZG_KEY_WINDOW) st = build_domain_tables(ctx, k);
    if (st == VKZG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = VKZG_ERR_CUDA;
    if (st != VKZG_OK) {
        cudaFree(k.bases);
        cudaFree(k.table);
        cudaFree(k.dom.omega);
        return st;
    }
    uint32_t id = ctx->next_key++;
    ctx->keys[id] = k;
    *key_id = id;
    return VKZG_OK;
}

int32_t vkzg_key_load(vkzg_ctx* ctx, const vkzg_g1_affine* bases, uint32_t n, const vkzg_g1_affine* q, uint32_t kind,
                      uint32_t window_bits, uint32_t* key_id) {
    VK_TRY(ctx_check(ctx));
    if (!bases || !n) return VKZG_ERR_ARG;
    DevBuf<affine_t> db, dq;
    VK_TRY(upload(ctx, db, bases, n));
    if (q) VK_TRY(upload(ctx, dq, q, 1));
    return vkzg_key_load_dev(ctx, (const vkzg_g1_affine*)db.p, n, q ? (const vkzg_g1_affine*)dq.p : nullptr, kind, window_bits,
                             key_id);
}

int32_t vkzg_key_free(vkzg_ctx* ctx, uint32_t key_id) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k) return VKZG_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(k->bases);
    cudaFree(k->table);
    cudaFree(k->dom.omega);
    ctx->keys.erase(key_id);
    return VKZG_OK;
}

uint64_t vkzg_key_table_bytes(const vkzg_ctx* ctx, uint32_t key_id) {
    if (!ctx) return 0;
    auto it = ctx->keys.find(key_id);
    return it == ctx->keys.end() ? 0 : it->second.table_points * sizeof(affine_t);
}

// ---- M1 ------------------------------------------------------------------------------------------------
int32_t vkzg_msm_range_dev(vkzg_ctx* ctx, uint32_t key_id, uint64_t first, const vkzg_fr* d_scalars, uint64_t n,
                           vkzg_g1_affine* d_out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_MSM || !d_out || (n && !d_scalars)) return VKZG_ERR_ARG;
    if (first > k->n || n > k->n - first) return VKZG_ERR_RANGE;  // (no wrap-around in first + n)
    return msm_large(ctx, *k, first, (const fp_t*)d_scalars, n, (affine_t*)d_out);
}
int32_t vkzg_msm_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_scalars, uint64_t n, vkzg_g1_affine* d_out) {
    return vkzg_msm_range_dev(ctx, key_id, 0, d_scalars, n, d_out);
}
int32_t vkzg_msm(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* scalars, uint64_t n, vkzg_g1_affine* out) {
    VK_TRY(ctx_check(ctx));
    if (!out || (n && !scalars)) return VKZG_ERR_ARG;
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_MSM) return VKZG_ERR_ARG;
    if (n > k->n) return VKZG_ERR_RANGE;
    DevBuf<fp_t> ds;
    DevBuf<affine_t> dout;
    VK_TRY(ds.alloc(ctx, n));
    VK_TRY(dout.alloc(ctx, 1));
    // the scalars cross PCIe in pieces and every piece is scattered into the bucket lists while the next one is in flight
    VK_TRY(msm_large(ctx, *k, 0, ds.p, n, dout.p, (const fp_t*)scalars));
    VK_TRY(download(ctx, out, dout.p, 1));
    return stream_sync(ctx);
}

int32_t vkzg_commit_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_scalars, uint32_t w, uint64_t B,
                              vkzg_g1_affine* d_out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || (B && (!d_scalars || !d_out))) return VKZG_ERR_ARG;
    if (w == 0 || w > k->n) return VKZG_ERR_RANGE;  // the reference zips against the key: at most n terms
    if (B == 0) return VKZG_OK;
    DevBuf<xyzz_t> acc;
    VK_TRY(acc.alloc(ctx, B));
    VK_TRY(fixed_base_msm(ctx, *k, (const fp_t*)d_scalars, w, B, 0, 0xffffffffu, acc));
    return normalize_points(ctx, acc, B, (affine_t*)d_out);
}
int32_t vkzg_commit_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* scalars, uint32_t w, uint64_t B, vkzg_g1_affine* out) {
    VK_TRY(ctx_check(ctx));
    if (B && (!scalars || !out)) return VKZG_ERR_ARG;
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (w == 0 || w > k->n) return VKZG_ERR_RANGE;
    if (B == 0) return VKZG_OK;
    DevBuf<fp_t> ds;
    DevBuf<affine_t> dout;
    DevBuf<xyzz_t> acc;
    VK_TRY(ds.alloc(ctx, (size_t)B * w));
    VK_TRY(dout.alloc(ctx, B));
    VK_TRY(acc.alloc(ctx, B));
    // chunks of the batch upload on the copy stream while the previous chunk is committed
    ChunkedUpload up(ctx);
    VK_TRY(up.init());
    for (uint64_t b0 = 0, nb = 0; b0 < B; b0 += nb) {
        nb = pipeline_piece(B, b0);
        VK_TRY(up.copy(ds.p + b0 * w, (const fp_t*)scalars + b0 * w, nb * w * sizeof(fp_t)));
        VK_TRY(up.publish());
        VK_TRY(fixed_base_msm(ctx, *k, ds.p + b0 * w, w, nb, 0, 0xffffffffu, acc.p + b0));
    }
    VK_TRY(normalize_points(ctx, acc, B, dout));
    VK_TRY(download(ctx, out, dout.p, B));
    return stream_sync(ctx);
}

int32_t vkzg_g1_sum_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_points, uint64_t n, vkzg_g1_affine* d_out) {
    VK_TRY(ctx_check(ctx));
    if (!d_out || (n && !d_points)) return VKZG_ERR_ARG;
    return g1_sum(ctx, (const affine_t*)d_points, n, (affine_t*)d_out);
}
int32_t vkzg_g1_sum(vkzg_ctx* ctx, const vkzg_g1_affine* points, uint64_t n, vkzg_g1_affine* out) {
    VK_TRY(ctx_check(ctx));
    if (!out || (n && !points)) return VKZG_ERR_ARG;
    DevBuf<affine_t> dp, dout;
    VK_TRY(upload(ctx, dp, points, n));
    VK_TRY(dout.alloc(ctx, 1));
    VK_TRY(g1_sum(ctx, dp, n, dout));
    VK_TRY(download(ctx, out, dout.p, 1));
    return stream_sync(ctx);
}

// ---- D1 ------------------------------------------------------------------------------------------------
int32_t vkzg_to_data_item_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_points, uint64_t n, vkzg_fr* d_out) {
    VK_TRY(ctx_check(ctx));
    if (n && (!d_points || !d_out)) return VKZG_ERR_ARG;
    return to_data_item(ctx, (const affine_t*)d_points, n, (fp_t*)d_out);
}
int32_t vkzg_to_data_item(vkzg_ctx* ctx, const vkzg_g1_affine* points, uint64_t n, vkzg_fr* out) {
    VK_TRY(ctx_check(ctx));
    if (n && (!points || !out)) return VKZG_ERR_ARG;
    DevBuf<affine_t> dp;
    DevBuf<fp_t> dout;
    VK_TRY(upload(ctx, dp, points, n));
    VK_TRY(dout.alloc(ctx, n));
    VK_TRY(to_data_item(ctx, dp, n, dout));
    VK_TRY(download(ctx, out, dout.p, n));
    return stream_sync(ctx);
}

}  // extern "C"

// ---- probes ---------------------------------------------------------------------------------------------
namespace vk {

__global__ void __launch_bounds__(256) k_probe_fq_mul(fp_t* x, const fp_t* y, uint64_t n, uint32_t iters) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t a = fp_load(x + i), b = fp_load(y + i);
#pragma unroll 1
    for (uint32_t k = 0; k < iters; ++k) a = fp_mul<Q>(a, b);
    fp_store(x + i, a);
}

// x <- x^(2^iters) through the hot loops' out-of-line multipliers: MODE 0 the dedicated square (fp_sqr_lazy), 1 the general
// lazy product on equal operands, 2 the Karatsuba product on equal operands; 3 / 4: x <- x b^iters (general / Karatsuba).  The device-side check of the square's carry chains and its throughput probe.
template <int MODE>
__global__ void __launch_bounds__(256) k_probe_fq_sqr(fp_t* x, uint64_t n, uint32_t iters) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t a = fp_load(x + i), b;
    // modes 3 / 4: a <- a * b with b = the operand's halves swapped (below 2^254)
#pragma unroll
    for (int k = 0; k < 8; ++k) b.l[k] = a.l[(k + 4) & 7];
    b.l[7] &= 0x3fffffffu;
#pragma unroll 1
    for (uint32_t k = 0; k < iters; ++k)
        a = MODE == 0 ? fp_sqr_lazy_ni<Q>(a) : MODE == 1 ? fp_mul_lazy_ni<Q>(a, a) : MODE == 2 ? fp_mul_lazy_kara_ni<Q>(a, a)
          : MODE == 3 ? fp_mul_lazy_ni<Q>(a, b) : fp_mul_lazy_kara_ni<Q>(a, b);
    fp_store(x + i, fp_canon<Q>(a));
}

// 8 independent accumulator chains per thread, 8 instructions per chain per iteration
template <int KIND>
__global__ void __launch_bounds__(256) k_probe_imad(uint32_t iters, uint64_t* sink) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = t * 2654435761u + 1, b = t ^ 0x9e3779b9u;
    if (KIND == 3) {
        double acc[8];
        double x = 1.0 + (double)(t & 1023) * 1e-9, y = 1.0 - (double)(t & 511) * 1e-9;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = (double)j;
#pragma unroll 1
        for (uint32_t k = 0; k < iters; ++k) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[j]) : "d"(x), "d"(y));
        }
        double s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[j];
        if (s == 12345.678) sink[0] = (uint64_t)s;
        return;
    }
    uint64_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = j + t;
#pragma unroll 1
    for (uint32_t k = 0; k < iters; ++k) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (KIND == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a), "r"(b));
            } else if (KIND == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t lo = (uint32_t)acc[j];
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
                    acc[j] = lo;
                }
            } else {
                // one carry chain across the 8 accumulators, like a Montgomery row
                uint32_t lo[8], hi[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    lo[j] = (uint32_t)acc[j];
                    hi[j] = (uint32_t)(acc[j] >> 32);
                }
                asm volatile(
                    "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
                    "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
                    "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
                    "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.cc.u32 %7, %16, %17, %7;\n\t"
                    "madc.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\t"
                    "madc.lo.cc.u32 %10, %16, %17, %10;\n\tmadc.hi.cc.u32 %11, %16, %17, %11;\n\t"
                    "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\t"
                    "madc.lo.cc.u32 %14, %16, %17, %14;\n\tmadc.hi.u32 %15, %16, %17, %15;"
                    : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
                      "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                    : "r"(a), "r"(b));
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = ((uint64_t)hi[j] << 32) | lo[j];
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= acc[j];
    if (s == 0x123456789abcdefull) sink[0] = s;
}

}  // namespace vk

extern "C" {

int32_t vkzg_probe_fq_mul_dev(vkzg_ctx* ctx, vkzg_fq* d_x, const vkzg_fq* d_y, uint64_t n, uint32_t iters) {
    VK_TRY(ctx_check(ctx));
    if (!d_x || !d_y || !n) return VKZG_ERR_ARG;
    k_probe_fq_mul<<<ceil_div_u64(n, 256), 256, 0, ctx->stream>>>((fp_t*)d_x, (const fp_t*)d_y, n, iters);
    return launch_check(ctx);
}

int32_t vkzg_probe_fq_sqr_dev(vkzg_ctx* ctx, vkzg_fq* d_x, uint64_t n, uint32_t iters, uint32_t mode) {
    VK_TRY(ctx_check(ctx));
    if (!d_x || !n || mode > 4) return VKZG_ERR_ARG;
    const uint32_t g = (uint32_t)ceil_div_u64(n, 256);
    switch (mode) {
        case 0: k_probe_fq_sqr<0><<<g, 256, 0, ctx->stream>>>((fp_t*)d_x, n, iters); break;
        case 1: k_probe_fq_sqr<1><<<g, 256, 0, ctx->stream>>>((fp_t*)d_x, n, iters); break;
        case 2: k_probe_fq_sqr<2><<<g, 256, 0, ctx->stream>>>((fp_t*)d_x, n, iters); break;
        case 3: k_probe_fq_sqr<3><<<g, 256, 0, ctx->stream>>>((fp_t*)d_x, n, iters); break;
        default: k_probe_fq_sqr<4><<<g, 256, 0, ctx->stream>>>((fp_t*)d_x, n, iters); break;
    }
    return launch_check(ctx);
}

int32_t vkzg_probe_imad_dev(vkzg_ctx* ctx, uint32_t kind, uint32_t blocks, uint32_t threads, uint32_t iters, uint64_t* macs_out) {
    VK_TRY(ctx_check(ctx));
    if (!blocks || !threads || threads > 256 || kind > 3) return VKZG_ERR_ARG;
    DevBuf<uint64_t> sink;
    VK_TRY(sink.alloc(ctx, 1));
    switch (kind) {
        case 0: k_probe_imad<0><<<blocks, threads, 0, ctx->stream>>>(iters, sink); break;
        case 1: k_probe_imad<1><<<blocks, threads, 0, ctx->stream>>>(iters, sink); break;
        case 2: k_probe_imad<2><<<blocks, threads, 0, ctx->stream>>>(iters, sink); break;
        default: k_probe_imad<3><<<blocks, threads, 0, ctx->stream>>>(iters, sink); break;
    }
    if (macs_out) *macs_out = (uint64_t)blocks * threads * iters * 64ull;
    return launch_check(ctx);
}

}  // extern "C"
