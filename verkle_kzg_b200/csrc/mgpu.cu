// Multi-GPU behind the C ABI: ONE host process drives several GPUs of a box — what a Rust
// caller of the vector-commit traits can use without torch.distributed.  A group owns one vkzg_ctx per device, each fed
// by its own host thread (the host-pointer entry points block until their result is in the caller's buffer).
//
//   width-N keys   replicated on every device; batches of commits / IPA proofs are cut into contiguous ranges, one per
//                  device, with no exchange (the path shards by independent vectors);
//   MSM keys       point-range sharded: device g holds the table of its slice; one MSM = per-device partial sums, each
//                  sent as a 64-byte affine point to device 0 over NVLink (cudaMemcpyPeerAsync, peer access enabled),
//                  added there by k_g1_sum.  Group addition is not an NCCL reduction operator and the payload is 64 bytes
//                  per device, so peer copies do what an all-gather would; the one-process-per-GPU form of the same
//                  exchange (torch.distributed all_gather over NCCL + vkzg_g1_sum_dev) is verkle_kzg_b200/sharding.py.
#include <thread>

#include "vk_common.cuh"

using namespace vk;

struct vkzg_mgpu {
    std::vector<vkzg_ctx*> ctx;
    std::vector<int> dev;
    struct MKey {
        uint32_t kind = 0;
        uint32_t n = 0;
        std::vector<uint32_t> id;       // per device
        std::vector<uint64_t> first;    // MSM keys: slice [first, first + count) of the points
        std::vector<uint64_t> count;
    };
    std::map<uint32_t, MKey> keys;
    uint32_t next_key = 1;
    affine_t* gather0 = nullptr;        // [ngpu] on device 0: the partial sums of one MSM
    std::vector<affine_t*> part;        // [1] on every device
};

namespace {

void split_range(uint64_t total, uint32_t world, uint32_t rank, uint64_t& first, uint64_t& count) {
    uint64_t base = total / world, rem = total % world;
    first = rank * base + (rank < rem ? rank : rem);
    count = base + (rank < rem ? 1 : 0);
}

// run f(g) for every device on its own host thread; first non-zero status wins
template <class F>
int32_t for_each_device(vkzg_mgpu* mg, F&& f) {
    const size_t n = mg->ctx.size();
    std::vector<int32_t> st(n, VKZG_OK);
    if (n == 1) return f(0);
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t g = 0; g < n; ++g) th.emplace_back([&, g] { st[g] = f((uint32_t)g); });
    for (auto& t : th) t.join();
    for (int32_t s : st)
        if (s != VKZG_OK) return s;
    return VKZG_OK;
}

}  // namespace

extern "C" {

int32_t vkzg_mgpu_create(vkzg_mgpu** out, const int32_t* device_ids, uint32_t ngpu) {
    if (!out) return VKZG_ERR_ARG;
    *out = nullptr;
    int count = 0;
    VK_CUDA(cudaGetDeviceCount(&count));
    if (ngpu == 0) {
        if (device_ids) return VKZG_ERR_ARG;
        ngpu = (uint32_t)count;
    }
    if (ngpu == 0 || ngpu > 64) return VKZG_ERR_ARG;
    auto* mg = new vkzg_mgpu();
    for (uint32_t g = 0; g < ngpu; ++g) {
        int d = device_ids ? device_ids[g] : (int)g;
        vkzg_ctx* c = nullptr;
        int32_t st = vkzg_ctx_create(&c, d);
        if (st != VKZG_OK) {
            for (auto* x : mg->ctx) vkzg_ctx_destroy(x);
            delete mg;
            return st;
        }
        mg->ctx.push_back(c);
        mg->dev.push_back(d);
    }
    // peer access towards device 0 (NVLink / NVSwitch): errors (same device, already enabled, no P2P) leave the staged
    // path of cudaMemcpyPeerAsync in place
    for (uint32_t g = 1; g < ngpu; ++g) {
        if (mg->dev[g] == mg->dev[0]) continue;
        cudaSetDevice(mg->dev[g]);
        if (cudaDeviceEnablePeerAccess(mg->dev[0], 0) != cudaSuccess) cudaGetLastError();
        cudaSetDevice(mg->dev[0]);
        if (cudaDeviceEnablePeerAccess(mg->dev[g], 0) != cudaSuccess) cudaGetLastError();
    }
    mg->part.resize(ngpu, nullptr);
    for (uint32_t g = 0; g < ngpu; ++g) {
        cudaSetDevice(mg->dev[g]);
        if (cudaMalloc((void**)&mg->part[g], sizeof(affine_t)) != cudaSuccess) {
            vkzg_mgpu_destroy(mg);
            return VKZG_ERR_OOM;
        }
    }
    cudaSetDevice(mg->dev[0]);
    if (cudaMalloc((void**)&mg->gather0, ngpu * sizeof(affine_t)) != cudaSuccess) {
        vkzg_mgpu_destroy(mg);
        return VKZG_ERR_OOM;
    }
    *out = mg;
    return VKZG_OK;
}

int32_t vkzg_mgpu_destroy(vkzg_mgpu* mg) {
    if (!mg) return VKZG_ERR_ARG;
    for (size_t g = 0; g < mg->ctx.size(); ++g) {
        cudaSetDevice(mg->dev[g]);
        if (g < mg->part.size() && mg->part[g]) cudaFree(mg->part[g]);
    }
    if (mg->gather0) {
        cudaSetDevice(mg->dev[0]);
        cudaFree(mg->gather0);
    }
    for (auto* c : mg->ctx) vkzg_ctx_destroy(c);
    delete mg;
    return VKZG_OK;
}

uint32_t vkzg_mgpu_size(const vkzg_mgpu* mg) { return mg ? (uint32_t)mg->ctx.size() : 0; }

vkzg_ctx* vkzg_mgpu_ctx(vkzg_mgpu* mg, uint32_t i) { return mg && i < mg->ctx.size() ? mg->ctx[i] : nullptr; }

int32_t vkzg_mgpu_key_load(vkzg_mgpu* mg, const vkzg_g1_affine* bases, uint32_t n, const vkzg_g1_affine* q, uint32_t kind,
                           uint32_t window_bits, uint32_t* key_id) {
    if (!mg || !bases || !n || !key_id) return VKZG_ERR_ARG;
    const uint32_t G = (uint32_t)mg->ctx.size();
    if (kind == VKZG_KEY_MSM && n < G) return VKZG_ERR_ARG;
    vkzg_mgpu::MKey k;
    k.kind = kind;
    k.n = n;
    k.id.assign(G, 0);
    k.first.assign(G, 0);
    k.count.assign(G, n);
    int32_t st = for_each_device(mg, [&](uint32_t g) -> int32_t {
        if (kind == VKZG_KEY_MSM) {
            split_range(n, G, g, k.first[g], k.count[g]);
            return vkzg_key_load(mg->ctx[g], bases + k.first[g], (uint32_t)k.count[g], nullptr, kind, window_bits, &k.id[g]);
        }
        return vkzg_key_load(mg->ctx[g], bases, n, q, kind, window_bits, &k.id[g]);
    });
    if (st != VKZG_OK) {
        for (uint32_t g = 0; g < G; ++g)
            if (k.id[g]) vkzg_key_free(mg->ctx[g], k.id[g]);
        return st;
    }
    uint32_t id = mg->next_key++;
    mg->keys[id] = k;
    *key_id = id;
    return VKZG_OK;
}

int32_t vkzg_mgpu_key_free(vkzg_mgpu* mg, uint32_t key_id) {
    if (!mg) return VKZG_ERR_ARG;
    auto it = mg->keys.find(key_id);
    if (it == mg->keys.end()) return VKZG_ERR_ARG;
    for (size_t g = 0; g < mg->ctx.size(); ++g) vkzg_key_free(mg->ctx[g], it->second.id[g]);
    mg->keys.erase(it);
    return VKZG_OK;
}

// one MSM over the first n points of a point-range sharded key (n <= key size: zip truncation, quirk Q1)
int32_t vkzg_mgpu_msm(vkzg_mgpu* mg, uint32_t key_id, const vkzg_fr* scalars, uint64_t n, vkzg_g1_affine* out) {
    if (!mg || !out || (n && !scalars)) return VKZG_ERR_ARG;
    auto it = mg->keys.find(key_id);
    if (it == mg->keys.end() || it->second.kind != VKZG_KEY_MSM) return VKZG_ERR_ARG;
    const vkzg_mgpu::MKey& k = it->second;
    if (n > k.n) return VKZG_ERR_RANGE;
    const uint32_t G = (uint32_t)mg->ctx.size();
    int32_t st = for_each_device(mg, [&](uint32_t g) -> int32_t {
        vkzg_ctx* c = mg->ctx[g];
        VK_TRY(ctx_check(c));
        // the part of [0, n) that falls into this device's slice
        uint64_t lo = k.first[g], hi = k.first[g] + k.count[g];
        uint64_t cnt = n > lo ? (n < hi ? n - lo : hi - lo) : 0;
        DevBuf<fp_t> ds;
        VK_TRY(upload(c, ds, (const fp_t*)scalars + lo, cnt));
        VK_TRY(vkzg_msm_dev(c, k.id[g], (const vkzg_fr*)ds.p, cnt, (vkzg_g1_affine*)mg->part[g]));
        // the 64-byte partial sum goes to device 0 over the peer link
        VK_CUDA(cudaMemcpyPeerAsync(mg->gather0 + g, mg->dev[0], mg->part[g], mg->dev[g], sizeof(affine_t), c->stream));
        return stream_sync(c);
    });
    VK_TRY(st);
    vkzg_ctx* c0 = mg->ctx[0];
    VK_TRY(ctx_check(c0));
    DevBuf<affine_t> dout;
    VK_TRY(dout.alloc(c0, 1));
    VK_TRY(g1_sum(c0, mg->gather0, G, dout));
    VK_TRY(download(c0, out, dout.p, 1));
    return stream_sync(c0);
}

// B independent commits of width w, contiguous batch ranges per device, no exchange
int32_t vkzg_mgpu_commit_batch(vkzg_mgpu* mg, uint32_t key_id, const vkzg_fr* scalars, uint32_t w, uint64_t B, vkzg_g1_affine* out) {
    if (!mg || (B && (!scalars || !out))) return VKZG_ERR_ARG;
    auto it = mg->keys.find(key_id);
    if (it == mg->keys.end() || it->second.kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    const vkzg_mgpu::MKey& k = it->second;
    const uint32_t G = (uint32_t)mg->ctx.size();
    return for_each_device(mg, [&](uint32_t g) -> int32_t {
        uint64_t first, cnt;
        split_range(B, G, g, first, cnt);
        if (cnt == 0) return w == 0 || w > k.n ? VKZG_ERR_RANGE : VKZG_OK;
        return vkzg_commit_batch(mg->ctx[g], k.id[g], scalars + first * w, w, cnt, out + first);
    });
}

// IPA::commit + IPA::prove_point over B vectors, contiguous batch ranges per device, no exchange
int32_t vkzg_mgpu_ipa_commit_prove_batch(vkzg_mgpu* mg, uint32_t key_id, const vkzg_fr* a, const vkzg_fr* points, uint64_t B,
                                         vkzg_g1_affine* commitments, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* y) {
    if (!mg || (B && (!a || !points || !commitments || !L || !R || !tip || !y))) return VKZG_ERR_ARG;
    auto it = mg->keys.find(key_id);
    if (it == mg->keys.end() || it->second.kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    const vkzg_mgpu::MKey& k = it->second;
    const uint32_t G = (uint32_t)mg->ctx.size();
    uint32_t rounds = 0;
    while ((1u << rounds) < k.n) ++rounds;
    return for_each_device(mg, [&](uint32_t g) -> int32_t {
        uint64_t first, cnt;
        split_range(B, G, g, first, cnt);
        if (cnt == 0) return VKZG_OK;
        return vkzg_ipa_commit_prove_batch(mg->ctx[g], k.id[g], a + first * k.n, points + first, cnt, commitments + first,
                                           L + first * rounds, R + first * rounds, tip + first, y + first);
    });
}

}  // extern "C"
