// Next row after the hot path: the verkle tree's host structure as the CALLER of the batched
// node commitment — VerkleTree::{new, insert_single, get_single, commitment} (verkle-tree/src/lib.rs:106-137) with
// Node::insert (node.rs:133-197) mirrored literally, including its quirks (the stem is the whole key; the internal
// node created on a stem collision is filed under the first differing unit but later indexed by tree depth).
// Keys are byte strings (Unit = u8, as every instantiation in the reference).
//
// Commitments are cached per node and cleared along the insertion path exactly like the reference
// (node.rs:41,46,145,156); vkzg_tree_commit recommits only the dirty nodes, level by level (k_tree_scalars +
// k_fixed_base_msm in CSR mode), feeding clean children in as already-known commitments.
#include <algorithm>
#include <array>

#include "vk_common.cuh"

namespace vk {

int32_t tree_level(vkzg_ctx* ctx, const Key& k, const uint32_t* d_row_ptr, uint64_t n_nodes, const uint16_t* d_slot,
                   const int32_t* d_child, const fp_t* d_lit, uint64_t n_terms, const affine_t* d_prev, affine_t* d_out, bool lit_raw);

struct HNode {
    bool internal = true;
    bool clean = false;          // commit below is valid
    affine_t commit;
    // internal
    std::vector<std::pair<uint8_t, uint32_t>> kids;  // (unit, node id), unsorted; direct table once it grows
    std::vector<int32_t> table;                      // 256 entries or empty
    // extension: the stem lives in the tree's byte arena (no per-node heap allocation), the first leaf inline
    uint64_t stem_off = 0;
    uint8_t leaf_unit = 0;
    std::array<uint8_t, 32> leaf_val;
};

}  // namespace vk

struct vkzg_tree {
    uint32_t key_len = 0, ext_width = 0;
    std::vector<vk::HNode> nodes;  // node 0 is the root
    std::vector<uint8_t> stems;    // key bytes of every extension node
    uint64_t n_keys = 0;
    uint64_t n_commits = 0;
    const uint8_t* stem(const vk::HNode& n) const { return stems.data() + n.stem_off; }

    int32_t child(uint32_t id, uint8_t unit) const {
        const vk::HNode& n = nodes[id];
        if (!n.table.empty()) return n.table[unit];
        for (auto& kv : n.kids)
            if (kv.first == unit) return (int32_t)kv.second;
        return -1;
    }
    void set_child(uint32_t id, uint8_t unit, uint32_t c) {
        vk::HNode& n = nodes[id];
        bool found = false;
        for (auto& kv : n.kids)
            if (kv.first == unit) {
                kv.second = c;
                found = true;
            }
        if (!found) n.kids.push_back({unit, c});
        if (!n.table.empty()) {
            n.table[unit] = (int32_t)c;
        } else if (n.kids.size() > 12) {
            n.table.assign(256, -1);
            for (auto& kv : n.kids) n.table[kv.first] = (int32_t)kv.second;
        }
    }
    uint32_t new_ext(const uint8_t* key, const uint8_t* value) {
        nodes.emplace_back();
        vk::HNode& e = nodes.back();
        e.internal = false;
        e.stem_off = stems.size();
        stems.insert(stems.end(), key, key + key_len);
        e.leaf_unit = key[key_len - 1];
        memcpy(e.leaf_val.data(), value, 32);
        return (uint32_t)nodes.size() - 1;
    }
    // Node::insert (node.rs:133-197).  Returns false where the reference panics (differing stem, node.rs:139-141).
    bool insert(const uint8_t* key, const uint8_t* value) {
        uint32_t cur = 0, depth = 0;
        for (;;) {
            nodes[cur].clean = false;  // cached commitments on the path are cleared
            if (depth >= key_len) return false;
            uint8_t k = key[depth];
            int32_t c = child(cur, k);
            if (c < 0) {
                uint32_t e = new_ext(key, value);
                set_child(cur, k, e);
                ++n_keys;
                return true;
            }
            if (!nodes[c].internal) {
                vk::HNode& ext = nodes[c];
                const uint8_t* est = stem(ext);
                bool same = memcmp(est, key, key_len) == 0;
                if (same || depth == key_len - 2) {
                    if (!same) return false;
                    ext.clean = false;
                    // same stem means same key, hence the same leaf unit: overwrite (node.rs:142-146)
                    memcpy(ext.leaf_val.data(), value, 32);
                    return true;
                }
                uint32_t d = depth + 1;  // next_diff_depth (lib.rs:50-59)
                while (d < key_len && est[d] == key[d]) ++d;
                if (d >= key_len) return false;  // the stems only differ above this depth: the reference indexes out of bounds
                uint8_t old_unit = est[d];
                nodes.emplace_back();
                uint32_t inner = (uint32_t)nodes.size() - 1;
                uint32_t e = new_ext(key, value);
                set_child(inner, key[d], e);
                set_child(inner, old_unit, (uint32_t)c);
                set_child(cur, k, inner);
                ++n_keys;
                return true;
            }
            cur = (uint32_t)c;
            ++depth;
        }
    }
    const uint8_t* get(const uint8_t* key) const {
        uint32_t cur = 0, depth = 0;
        while (nodes[cur].internal) {
            if (depth >= key_len) return nullptr;
            int32_t c = child(cur, key[depth]);
            if (c < 0) return nullptr;
            cur = (uint32_t)c;
            ++depth;
        }
        const vk::HNode& e = nodes[cur];
        if (memcmp(stem(e), key, key_len) != 0) return nullptr;
        return e.leaf_unit == key[key_len - 1] ? e.leaf_val.data() : nullptr;
    }
};

namespace vk {

// from_le_bytes_mod_order of a byte string -> Montgomery Fr (host_hash.cpp, native 64-bit limbs)
extern "C" void vkh_fr_from_le_bytes(const uint8_t* b, size_t len, uint64_t out[4]);
static fp_t fr_from_le_bytes(const uint8_t* b, size_t len) {
    fp_t r;
    vkh_fr_from_le_bytes(b, len, (uint64_t*)r.l);
    return r;
}

struct LevelBuf {
    std::vector<uint32_t> row_ptr{0};
    std::vector<uint16_t> slot;
    std::vector<int32_t> child;
    std::vector<fp_t> lit;
    std::vector<uint32_t> owner;  // node id per row (UINT32_MAX for the C1 / C2 helper rows)
    void term(uint16_t s, int32_t c, const fp_t& l) {
        slot.push_back(s);
        child.push_back(c);
        lit.push_back(l);
    }
    uint32_t close(uint32_t node) {
        row_ptr.push_back((uint32_t)slot.size());
        owner.push_back(node);
        return (uint32_t)owner.size() - 1;
    }
};

}  // namespace vk

using namespace vk;

extern "C" {

int32_t vkzg_tree_create(vkzg_tree** out, uint32_t key_len, uint32_t ext_width) {
    if (!out || key_len < 2 || ext_width == 0 || ext_width > 65535) return VKZG_ERR_ARG;
    vkzg_tree* t = new vkzg_tree();
    t->key_len = key_len;
    t->ext_width = ext_width;
    t->nodes.emplace_back();  // root: Internal, no children
    *out = t;
    return VKZG_OK;
}

int32_t vkzg_tree_destroy(vkzg_tree* t) {
    delete t;
    return VKZG_OK;
}

// n (key, value) pairs inserted IN ORDER; stops at the first pair the reference would panic on (VKZG_ERR_RANGE,
// *n_done = pairs inserted)
int32_t vkzg_tree_insert(vkzg_tree* t, const uint8_t* keys, const uint8_t* values, uint64_t n, uint64_t* n_done) {
    if (!t || (n && (!keys || !values))) return VKZG_ERR_ARG;
    t->nodes.reserve(t->nodes.size() + n + n / 4 + 16);
    t->stems.reserve(t->stems.size() + n * t->key_len);
    for (uint64_t i = 0; i < n; ++i) {
        if (!t->insert(keys + i * t->key_len, values + i * 32)) {
            if (n_done) *n_done = i;
            return VKZG_ERR_RANGE;
        }
    }
    if (n_done) *n_done = n;
    return VKZG_OK;
}

// returns 1 and copies the 32-byte value if present, 0 otherwise
int32_t vkzg_tree_get(const vkzg_tree* t, const uint8_t* key, uint8_t* value_out) {
    if (!t || !key) return VKZG_ERR_ARG;
    const uint8_t* v = t->get(key);
    if (!v) return 0;
    if (value_out) memcpy(value_out, v, 32);
    return 1;
}

uint64_t vkzg_tree_nodes(const vkzg_tree* t) { return t ? t->nodes.size() : 0; }

// VerkleTree::commitment (lib.rs:127-129): recommit every dirty node, leaves first, return the root commitment.
// *n_committed (optional) = number of node commitments computed by this call (C1 / C2 helper vectors included).
int32_t vkzg_tree_commit(vkzg_ctx* ctx, uint32_t key_id, vkzg_tree* t, vkzg_g1_affine* root_out, uint64_t* n_committed) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || !t || !root_out) return VKZG_ERR_ARG;
    if (k->n < 256 || k->n < t->ext_width || k->n < 4) return VKZG_ERR_RANGE;
    if (n_committed) *n_committed = 0;
    if (t->nodes[0].clean) {
        memcpy(root_out, &t->nodes[0].commit, sizeof(affine_t));
        return VKZG_OK;
    }
    const uint32_t W = t->ext_width;
    // literals travel as raw little-endian integers when they fit 32 bytes (the device reduces and converts them);
    // longer stems are reduced on the host and everything travels in Montgomery form
    const bool raw = t->key_len <= 32;
    auto literal = [raw](const uint8_t* b, size_t len) {
        if (!raw) return fr_from_le_bytes(b, len);
        fp_t v = fp_zero<S>();
        memcpy(v.l, b, len);
        return v;
    };
    // ---- post-order over the dirty part: heights -> levels.  Level 0 = C1 / C2 helper vectors, 1 = extensions,
    //      >= 2 internal nodes by height.  Clean children enter as known commitments (prefix of the node array).
    std::vector<LevelBuf> levels(2);
    if (t->n_commits == 0) {  // first (bulk) commit: every node is dirty, size the leaf-side levels once
        const size_t nn = t->nodes.size();
        levels[0].slot.reserve(2 * nn); levels[0].child.reserve(2 * nn); levels[0].lit.reserve(2 * nn);
        levels[0].row_ptr.reserve(nn + 1); levels[0].owner.reserve(nn);
        levels[1].slot.reserve(3 * nn); levels[1].child.reserve(3 * nn); levels[1].lit.reserve(3 * nn);
        levels[1].row_ptr.reserve(nn + 1); levels[1].owner.reserve(nn);
    }
    ++t->n_commits;
    std::vector<affine_t> known;            // commitments of clean children referenced by dirty parents
    std::vector<std::pair<uint32_t, int32_t>> handle(t->nodes.size(), {0xffffffffu, -1});  // node -> (level, row) ; level 0xfffffffe = known
    std::vector<std::pair<uint32_t, bool>> stack{{0u, false}};
    fp_t one = fp_one<S>(), zero = fp_zero<S>();
    if (raw) {
        one = zero;
        one.l[0] = 1;
    }
    while (!stack.empty()) {
        auto [id, done] = stack.back();
        stack.pop_back();
        HNode& n = t->nodes[id];
        if (n.clean) {
            known.push_back(n.commit);
            handle[id] = {0xfffffffeu, (int32_t)known.size() - 1};
            continue;
        }
        if (!n.internal) {
            // node.rs:226-240: leaf idx -> slots (2 idx) % W, (2 idx + 1) % W of C1 (idx < W/2) or C2; later leaves overwrite
            // one leaf per extension (the stem is the whole key): its two halves go to C1 or to C2
            int32_t r1 = -1, r2 = -1;
            {
                uint32_t idx = n.leaf_unit;
                uint16_t s_lo = (uint16_t)((2 * idx) % W), s_hi = (uint16_t)((2 * idx + 1) % W);
                if (s_lo != s_hi) levels[0].term(s_lo, -1, literal(n.leaf_val.data(), 16));  // W == 1: the later write (high) wins
                levels[0].term(s_hi, -1, literal(n.leaf_val.data() + 16, 16));
                int32_t row = (int32_t)levels[0].close(0xffffffffu);
                if (idx < W / 2)
                    r1 = row;
                else
                    r2 = row;
            }
            // node.rs:243-253: commit([1, stem, C1, C2]); child ids are patched to global ids below (level 0 rows)
            levels[1].term(0, -1, one);
            levels[1].term(1, -1, literal(t->stem(n), t->key_len));
            if (r1 >= 0) levels[1].term(2, -(r1 + 2), zero);  // encoded: -(row + 2) = level-0 row, resolved after layout
            if (r2 >= 0) levels[1].term(3, -(r2 + 2), zero);
            handle[id] = {1u, (int32_t)levels[1].close(id)};
            continue;
        }
        if (!done) {
            stack.push_back({id, true});
            for (auto& kv : n.kids) stack.push_back({kv.second, false});
            continue;
        }
        uint32_t level = 2;
        for (auto& kv : n.kids) {
            auto h = handle[kv.second];
            if (h.first != 0xfffffffeu) level = std::max(level, h.first + 1);
        }
        if (levels.size() <= level) levels.resize(level + 1);
        std::vector<std::pair<uint8_t, uint32_t>> kids = n.kids;
        std::sort(kids.begin(), kids.end());
        for (auto& kv : kids) {
            // child reference encoded as (level << 40 | row) is too wide for int32: store node id, resolve below
            levels[level].term(kv.first, (int32_t)kv.second, zero);
        }
        handle[id] = {level, (int32_t)levels[level].close(id)};
    }
    // ---- internal levels: order the nodes of a level by their number of children, so that the lanes / groups of one
    //      warp (one node each) walk term lists of similar length instead of waiting for the longest
    for (size_t l = 2; l < levels.size(); ++l) {
        LevelBuf& L = levels[l];
        const size_t nn = L.owner.size();
        if (nn < 64) continue;
        std::vector<uint32_t> perm(nn);
        for (size_t j = 0; j < nn; ++j) perm[j] = (uint32_t)j;
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) {
            return L.row_ptr[a + 1] - L.row_ptr[a] > L.row_ptr[b + 1] - L.row_ptr[b];
        });
        LevelBuf S;
        S.slot.reserve(L.slot.size());
        S.child.reserve(L.child.size());
        S.lit.reserve(L.lit.size());
        for (size_t j = 0; j < nn; ++j) {
            uint32_t r = perm[j];
            for (uint32_t t2 = L.row_ptr[r]; t2 < L.row_ptr[r + 1]; ++t2) S.term(L.slot[t2], L.child[t2], L.lit[t2]);
            uint32_t row = S.close(L.owner[r]);
            handle[L.owner[r]] = {(uint32_t)l, (int32_t)row};
        }
        L = std::move(S);
    }
    // ---- global ids: [known commitments][level 0][level 1]...
    std::vector<uint64_t> base(levels.size(), 0);
    uint64_t total = known.size();
    for (size_t l = 0; l < levels.size(); ++l) {
        base[l] = total;
        total += levels[l].owner.size();
    }
    if (total >= (1ull << 31)) return VKZG_ERR_RANGE;
    for (auto& c : levels[1].child)
        if (c <= -2) c = (int32_t)(base[0] + (uint64_t)(-c - 2));
    for (size_t l = 2; l < levels.size(); ++l)
        for (auto& c : levels[l].child) {
            auto h = handle[(uint32_t)c];
            c = h.first == 0xfffffffeu ? h.second : (int32_t)(base[h.first] + (uint64_t)h.second);
        }
    // ---- device passes
    DevBuf<affine_t> all;
    VK_TRY(all.alloc(ctx, total));
    if (!known.empty()) VK_CUDA(cudaMemcpyAsync(all.p, known.data(), known.size() * sizeof(affine_t), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t l = 0; l < levels.size(); ++l) {
        LevelBuf& L = levels[l];
        uint64_t nn = L.owner.size();
        if (!nn) continue;
        DevBuf<uint32_t> d_rp;
        DevBuf<uint16_t> d_sl;
        DevBuf<int32_t> d_ch;
        DevBuf<fp_t> d_li;
        VK_TRY(upload(ctx, d_rp, L.row_ptr.data(), nn + 1));
        VK_TRY(upload(ctx, d_sl, L.slot.data(), L.slot.size()));
        VK_TRY(upload(ctx, d_ch, L.child.data(), L.child.size()));
        VK_TRY(upload(ctx, d_li, L.lit.data(), L.lit.size()));
        VK_TRY(tree_level(ctx, *k, d_rp, nn, d_sl, d_ch, d_li, L.slot.size(), all.p, all.p + base[l], raw));
        VK_TRY(stream_sync(ctx));  // the host vectors of this level go out of use before the next upload reuses the pool
    }
    // ---- cache the new commitments on the host
    std::vector<affine_t> host(total - known.size());
    if (!host.empty()) VK_CUDA(cudaMemcpyAsync(host.data(), all.p + known.size(), host.size() * sizeof(affine_t), cudaMemcpyDeviceToHost, ctx->stream));
    VK_TRY(stream_sync(ctx));
    uint64_t off = 0;
    for (size_t l = 0; l < levels.size(); ++l) {
        for (size_t j = 0; j < levels[l].owner.size(); ++j) {
            uint32_t id = levels[l].owner[j];
            if (id != 0xffffffffu) {
                t->nodes[id].commit = host[off + j];
                t->nodes[id].clean = true;
            }
        }
        off += levels[l].owner.size();
    }
    if (n_committed) *n_committed = host.size();
    memcpy(root_out, &t->nodes[0].commit, sizeof(affine_t));
    return VKZG_OK;
}

}  // extern "C"
