// Shared host-side plumbing of libvkzg: context, keys, stream-ordered scratch memory, launch counting.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "../../include/vkzg.h"
#include "hash.cuh"

namespace vk {

#define VK_CUDA(expr)                                                                                          \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            fprintf(stderr, "[vkzg] %s failed at %s:%d: %s\n", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return _e == cudaErrorMemoryAllocation ? VKZG_ERR_OOM : VKZG_ERR_CUDA;                              \
        }                                                                                                      \
    } while (0)

#define VK_TRY(expr)                 \
    do {                             \
        int32_t _s = (expr);         \
        if (_s != VKZG_OK) return _s; \
    } while (0)

static_assert(sizeof(fp_t) == sizeof(vkzg_fr), "layout");
static_assert(sizeof(affine_t) == sizeof(vkzg_g1_affine), "layout");

static const int WARPS_PER_CTA = 4;
static const int CHUNK_TERMS = 128;  // terms recoded per shared-memory chunk in the fixed-base kernel

// Domain constants of a width-N key (PrecomputedLagrange, precompute.rs:11-34), built on the device at
// key load: omega^i, omega^-i, 1/(omega^d - 1), N^-1.
struct DomainTables {
    fp_t* omega = nullptr;       // [N]  w^i
    fp_t* omega_inv = nullptr;   // [N]  w^-i
    fp_t* diff_inv = nullptr;    // [N]  1/(w^d - 1), entry 0 unused (0)
    fp_t n_inv;                  // 1/N
    fp_t n_mont;                 // N
};

struct Key {
    uint32_t kind = 0;
    uint32_t n = 0;          // number of g bases
    bool has_q = false;      // base index n is q
    uint32_t c = 0;          // window bits
    uint32_t W = 0;          // windows
    affine_t* bases = nullptr;  // [n (+1)] plain copy
    affine_t* table = nullptr;  // WINDOW: [(base*W + w) << (c-1) | mag-1] ; MSM: [w*n + i]
    uint64_t table_points = 0;
    uint32_t log2n = 0;      // WINDOW keys: domain = next pow2(n)
    uint32_t domain_n = 0;
    DomainTables dom;
};

}  // namespace vk

struct vkzg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    int sm_count = 0;
    uint32_t next_key = 1;
    std::map<uint32_t, vk::Key> keys;
    // optional per-kernel timing of the dominant kernel (bench.py's roofline): CUDA event pairs around
    // every k_fixed_base_msm / k_msm_bucket launch on this context's stream
    cudaStream_t copy_stream = nullptr;  // host<->device staging of the batched host-pointer calls overlaps compute
    cudaStream_t aux_stream = nullptr;   // second compute stream: two half-batches of IPA proofs run interleaved
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // its fork / join events (created once with the stream)
    std::vector<cudaStream_t> side_streams;  // batched multiproofs: the per-proof row kernels of different proofs overlap
    bool ipa_two_streams = true;         // VKZG_OPT_IPA_TWO_STREAMS
    int batch_affine = -1;               // VKZG_OPT_BATCH_AFFINE: -1 automatic, 0 off, 1 on (big dense batches)
    bool multiproof_check_y = false;     // VKZG_OPT_MULTIPROOF_CHECK_Y (diagnostic, see include/vkzg.h)
    cudaMemPool_t pool = nullptr;        // private stream-ordered scratch pool (api.cu: ctx_create)
    int tree_flatten = 0;                // VKZG_OPT_TREE_FLATTEN
    // grow-only pinned host staging areas (vkzg_tree_commit: compact node records up, commitments down); pages stay
    // resident and DMA-able between calls instead of being faulted in and staged by the driver every time
    void* host_stage[2] = {nullptr, nullptr};
    size_t host_stage_bytes[2] = {0, 0};
    void* pinned_stage(int slot, size_t bytes) {
        if (bytes > host_stage_bytes[slot]) {
            if (host_stage[slot]) cudaFreeHost(host_stage[slot]);
            host_stage[slot] = nullptr;
            host_stage_bytes[slot] = 0;
            size_t want = bytes + bytes / 4;
            if (cudaMallocHost(&host_stage[slot], want) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            host_stage_bytes[slot] = want;
        }
        return host_stage[slot];
    }
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    uint64_t timing_units = 0;  // point additions issued by the timed launches (upper bound: zero digits excluded at run time)
    std::map<uint32_t, vk::DomainTables> domains;  // data-domain constants by log2(size), built on demand

    vk::Key* key(uint32_t id) {
        auto it = keys.find(id);
        return it == keys.end() ? nullptr : &it->second;
    }
};

namespace vk {

// stream-ordered scratch buffer (cudaMallocAsync pool keeps the memory cached between calls)
template <class T>
struct DevBuf {
    T* p = nullptr;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    int32_t alloc(vkzg_ctx* ctx, size_t count) {
        s = ctx->stream;
        if (count == 0) count = 1;
        cudaError_t e = cudaMallocFromPoolAsync((void**)&p, count * sizeof(T), ctx->pool, s);
        if (e != cudaSuccess) {
            fprintf(stderr, "[vkzg] cudaMallocFromPoolAsync(%zu bytes) failed: %s\n", count * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            p = nullptr;
            return VKZG_ERR_OOM;
        }
        return VKZG_OK;
    }
    ~DevBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    operator T*() const { return p; }
};

static inline int32_t launch_check(vkzg_ctx* ctx, int n = 1) {
    ctx->launches += n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "[vkzg] kernel launch failed: %s\n", cudaGetErrorString(e));
        return VKZG_ERR_CUDA;
    }
    return VKZG_OK;
}

// bracket a launch of the dominant kernel with an event pair when timing is on
struct KernelTimer {
    vkzg_ctx* ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    explicit KernelTimer(vkzg_ctx* c) : ctx(c) {
        if (ctx->timing && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, ctx->stream);
    }
    ~KernelTimer() {
        if (a && b) {
            cudaEventRecord(b, ctx->stream);
            ctx->timing_events.emplace_back(a, b);
        }
    }
};

static inline int32_t ctx_check(vkzg_ctx* ctx) {
    if (!ctx) return VKZG_ERR_ARG;
    VK_CUDA(cudaSetDevice(ctx->device));
    return VKZG_OK;
}

// host <-> device staging for the host-pointer entry points
template <class T>
static inline int32_t upload(vkzg_ctx* ctx, DevBuf<T>& buf, const void* host, size_t count) {
    VK_TRY(buf.alloc(ctx, count));
    if (count) VK_CUDA(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return VKZG_OK;
}
template <class T>
static inline int32_t download(vkzg_ctx* ctx, void* host, const T* dev, size_t count) {
    if (count) VK_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return VKZG_OK;
}
static inline int32_t stream_sync(vkzg_ctx* ctx) {
    VK_CUDA(cudaStreamSynchronize(ctx->stream));
    return VKZG_OK;
}

// Chunked upload for the batched host-pointer entry points: the rows of chunk i + 1 cross PCIe on the copy stream
// while chunk i computes on the main stream (pinned host memory; pageable memory degrades to synchronous staging).
struct ChunkedUpload {
    vkzg_ctx* ctx;
    std::vector<cudaEvent_t> evs;
    explicit ChunkedUpload(vkzg_ctx* c) : ctx(c) {}
    ~ChunkedUpload() {
        for (auto e : evs) cudaEventDestroy(e);
    }
    int32_t init() {
        if (!ctx->copy_stream) VK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        // the destination buffers were allocated in main-stream order: the copy stream must not run ahead of that
        cudaEvent_t e;
        VK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        evs.push_back(e);
        VK_CUDA(cudaEventRecord(e, ctx->stream));
        VK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, e, 0));
        return VKZG_OK;
    }
    // enqueue dst <- src on the copy stream
    int32_t copy(void* dst, const void* src, size_t bytes) {
        if (bytes) VK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        return VKZG_OK;
    }
    // main stream waits for everything enqueued on the copy stream so far
    int32_t publish() {
        cudaEvent_t e;
        VK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        evs.push_back(e);
        VK_CUDA(cudaEventRecord(e, ctx->copy_stream));
        VK_CUDA(cudaStreamWaitEvent(ctx->stream, e, 0));
        return VKZG_OK;
    }
};

// Pieces of a pipelined batch upload (piece k is committed while piece k + 1 crosses PCIe).  The first piece is what the GPU
// waits for and every piece is a kernel launch with its own partial last wave, so the schedule is geometric — B/16, 3B/16, the
// rest in ONE launch: a width-256 commit computes ~3.5x longer than its scalars take to arrive, each piece's upload hides under
// the piece before it (2^14 commits from host buffers on one GPU: 1.756 M/s against 1.677 M/s with four equal pieces; what pays is
// the single big last launch — B/16, 3B/16 followed by quarters measured 1.689 M/s).  The price: when eight ranks share the host's
// PCIe and memory the 3/4 piece no longer hides (8 x 2^14 commits from host buffers 11.1 M/s against 12.1 M/s) —
// VKZG_PIPE_UNIFORM=1 restores four equal pieces for such callers.
static inline uint64_t pipeline_piece(uint64_t B, uint64_t b0) {
    if (B < 8192) return B - b0;
    static int uniform = -1;
    if (uniform < 0) uniform = getenv("VKZG_PIPE_UNIFORM") ? 1 : 0;
    uint64_t n;
    if (uniform) {
        n = (B + 3) / 4;
        if (n < 4096) n = 4096;
    } else {
        const uint64_t first = B / 16, second = 3 * B / 16;
        n = b0 == 0 ? first : (b0 == first ? second : B - b0);
    }
    return n < B - b0 ? n : B - b0;
}

static inline uint32_t ceil_div_u64(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
// dynamic shared memory of one k_fixed_base_msm CTA (commit.cu): per warp a list of CHUNK_TERMS * W entries + 32 counters
static inline size_t fixed_base_smem_bytes(uint32_t W) { return (size_t)WARPS_PER_CTA * (CHUNK_TERMS * W + 32) * sizeof(uint32_t); }

// ---- internal entry points implemented across the .cu files (all take device pointers) -------------
int32_t normalize_points(vkzg_ctx* ctx, const xyzz_t* d_in, uint64_t n, affine_t* d_out);
// VKZG_ERR_ARG unless every point is (0,0) or a canonical point with y^2 = x^3 + 3 (synchronises the stream)
int32_t check_points_on_curve(vkzg_ctx* ctx, const affine_t* d_points, uint64_t n);
int32_t build_window_tables(vkzg_ctx* ctx, Key& k);
int32_t build_msm_tables(vkzg_ctx* ctx, Key& k);
int32_t build_domain_tables(vkzg_ctx* ctx, Key& k);
int32_t domain_for(vkzg_ctx* ctx, uint32_t lg, const DomainTables*& dt);
// jobs x T-term fixed-base MSMs.  scalars[jobs][T] (Montgomery Fr).  ipa_m == 0: term t uses base t.
// ipa_m != 0: the L/R cross-term base selection of an IPA round (commit.cu).  Result: out_xyzz[jobs].
int32_t fixed_base_msm(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                       uint32_t q_row, xyzz_t* d_out);
// the same through the batch-affine tree (batch_affine.cu): dense jobs only, pays from a few thousand jobs on
int32_t fixed_base_msm_batch_affine(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                                    uint32_t q_row, xyzz_t* d_out);
// CSR variant (verkle nodes): job j owns terms [row_ptr[j], row_ptr[j+1]), term t uses base slot[t]
int32_t fixed_base_msm_csr(vkzg_ctx* ctx, const Key& k, const fp_t* d_scalars, uint32_t T, uint64_t jobs, uint32_t ipa_m,
                           uint32_t q_row, const uint32_t* d_row_ptr, const uint16_t* d_slot, xyzz_t* d_out, uint32_t lanes_per_job,
                           uint32_t csr_split);
int32_t barycentric_batch(vkzg_ctx* ctx, const Key& k, const fp_t* d_points, uint64_t B, fp_t* d_out);
int32_t ipa_prove_core(vkzg_ctx* ctx, const Key& k, int mode, uint32_t N, const fp_t* d_a, const fp_t* d_points,
                       const affine_t* d_C, uint64_t B, const uint8_t* prefix, uint32_t prefix_len, const char* dst, affine_t* d_L,
                       affine_t* d_R, fp_t* d_tip, fp_t* d_y, const uint8_t* d_prefix_each = nullptr, uint32_t each_len = 0);
int32_t ipa_verify_core(vkzg_ctx* ctx, const Key& k, int mode, const fp_t* d_points, const affine_t* d_C, uint64_t B, const uint8_t* prefix,
                        uint32_t prefix_len, const char* dst, const affine_t* d_L, const affine_t* d_R, const fp_t* d_tip,
                        const fp_t* d_y, int32_t* d_ok);
int32_t poly_batch(vkzg_ctx* ctx, const Key& k, const fp_t* d_f, uint32_t len, uint32_t domain_n, const fp_t* d_points, uint64_t B,
                   fp_t* d_q, fp_t* d_y, bool check_err, uint32_t share = 1);
int32_t kzg_open_core(vkzg_ctx* ctx, const Key& k, const fp_t* d_f, uint32_t len, uint32_t domain_n, const fp_t* d_points,
                      uint64_t B, affine_t* d_proof, fp_t* d_y, bool check_err);
int32_t var_base_msm(vkzg_ctx* ctx, const affine_t* d_points, const fp_t* d_scalars, uint64_t n, affine_t* d_out);
int32_t msm_large(vkzg_ctx* ctx, const Key& k, uint64_t first, const fp_t* d_scalars, uint64_t n, affine_t* d_out,
                  const fp_t* h_scalars = nullptr);
int32_t g1_sum(vkzg_ctx* ctx, const affine_t* d_points, uint64_t n, affine_t* d_out);
int32_t to_data_item(vkzg_ctx* ctx, const affine_t* d_points, uint64_t n, fp_t* d_out);

}  // namespace vk
