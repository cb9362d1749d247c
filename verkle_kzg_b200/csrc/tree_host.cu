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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "vk_common.cuh"

namespace vk {

int32_t tree_level(vkzg_ctx* ctx, const Key& k, const uint32_t* d_row_ptr, uint64_t n_nodes, const uint16_t* d_slot,
                   const int32_t* d_child, const fp_t* d_lit, uint64_t n_terms, const affine_t* d_prev, affine_t* d_out, bool lit_raw);
int32_t tree_ext_levels(vkzg_ctx* ctx, const Key& k, const fp_t* d_stem, const uint8_t* d_unit, const uint8_t* d_val, uint64_t n,
                        uint32_t W, affine_t* d_all, uint64_t base0, uint64_t base1);

struct HNode {
    static constexpr uint32_t INLINE = 12;  // children kept inline before the node switches to a direct table
    bool internal = true;
    uint8_t nk = 0;              // inline children in use (internal nodes without a table)
    uint8_t leaf_unit = 0;       // extension: the first leaf lives inline
    int32_t table_id = -1;       // internal: index of this node's 256-entry direct table in the tree's table arena
    uint8_t unit[INLINE];        // internal: (unit, node id) pairs in insertion order
    uint32_t kid[INLINE];
    uint64_t stem_off = 0;       // extension: the stem lives in the tree's byte arena (no per-node heap allocation)
    std::array<uint8_t, 32> leaf_val;
};

}  // namespace vk

struct vkzg_tree {
    uint32_t key_len = 0, ext_width = 0;
    std::vector<vk::HNode> nodes;  // node 0 is the root
    std::vector<uint8_t> stems;    // key bytes of every extension node
    std::vector<int32_t> tables;   // 256 child ids (-1 = none) per internal node that outgrew its inline list
    std::vector<vk::affine_t> commits;  // cached commitment per node id (valid where clean[id]); sized at commit time
    std::vector<uint8_t> clean;    // per node id: the cached commitment is valid.  Kept out of the node records: the flags of a
                                   // million nodes stay cache-resident for insertion (cleared along the path, node.rs:41,46,145,156)
                                   // and for the sequential passes of vkzg_tree_commit
    bool is_clean(uint32_t id) const { return id < clean.size() && clean[id]; }
    uint64_t n_dirty = 1;          // nodes whose cached commitment is missing or cleared (the empty root to begin with)
    uint64_t n_keys = 0;
    uint64_t n_commits = 0;
    const uint8_t* stem(const vk::HNode& n) const { return stems.data() + n.stem_off; }
    const int32_t* table(const vk::HNode& n) const { return tables.data() + (size_t)n.table_id * 256; }

    int32_t child(uint32_t id, uint8_t unit) const {
        const vk::HNode& n = nodes[id];
        if (n.table_id >= 0) return table(n)[unit];
        for (uint32_t j = 0; j < n.nk; ++j)
            if (n.unit[j] == unit) return (int32_t)n.kid[j];
        return -1;
    }
    void set_child(uint32_t id, uint8_t unit, uint32_t c) {
        vk::HNode& n = nodes[id];
        if (n.table_id >= 0) {
            tables[(size_t)n.table_id * 256 + unit] = (int32_t)c;
            return;
        }
        for (uint32_t j = 0; j < n.nk; ++j)
            if (n.unit[j] == unit) {
                n.kid[j] = c;
                return;
            }
        if (n.nk < vk::HNode::INLINE) {
            n.unit[n.nk] = unit;
            n.kid[n.nk] = c;
            ++n.nk;
            return;
        }
        n.table_id = (int32_t)(tables.size() / 256);
        tables.resize(tables.size() + 256, -1);
        int32_t* tb = tables.data() + (size_t)n.table_id * 256;
        for (uint32_t j = 0; j < n.nk; ++j) tb[n.unit[j]] = (int32_t)n.kid[j];
        tb[unit] = (int32_t)c;
    }
    // children in ascending unit order
    template <typename F>
    void for_each_child(const vk::HNode& n, F&& f) const {
        if (n.table_id >= 0) {
            const int32_t* tb = table(n);
            for (uint32_t u = 0; u < 256; ++u)
                if (tb[u] >= 0) f((uint8_t)u, (uint32_t)tb[u]);
            return;
        }
        uint8_t order[vk::HNode::INLINE];
        for (uint32_t j = 0; j < n.nk; ++j) order[j] = (uint8_t)j;
        std::sort(order, order + n.nk, [&](uint8_t a, uint8_t b) { return n.unit[a] < n.unit[b]; });
        for (uint32_t j = 0; j < n.nk; ++j) f(n.unit[order[j]], n.kid[order[j]]);
    }
    uint32_t new_ext(const uint8_t* key, const uint8_t* value) {
        ++n_dirty;
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
            if (is_clean(cur)) {  // cached commitments on the path are cleared
                clean[cur] = 0;
                ++n_dirty;
            }
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
                    if (is_clean((uint32_t)c)) {
                        clean[c] = 0;
                        ++n_dirty;
                    }
                    // same stem means same key, hence the same leaf unit: overwrite (node.rs:142-146)
                    memcpy(ext.leaf_val.data(), value, 32);
                    return true;
                }
                uint32_t d = depth + 1;  // next_diff_depth (lib.rs:50-59)
                while (d < key_len && est[d] == key[d]) ++d;
                if (d >= key_len) return false;  // the stems only differ above this depth: the reference indexes out of bounds
                uint8_t old_unit = est[d];
                ++n_dirty;
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
    // Software prefetch for bulk insertion.  A walk along `key` is a chain of dependent reads: node header -> child slot
    // (direct table entry; small nodes keep their children inline) -> next node header -> ... -> stem bytes.  prefetch_path performs the
    // first `reads` of them on the CURRENT structure (read-only; the structure may still change before the key is
    // inserted, a stale prefetch is harmless) and prefetches the target of the next one WITHOUT touching it.  The bulk
    // loop calls it with increasing `reads` at decreasing look-ahead distances, so each call only dereferences lines
    // an earlier call requested.
    void prefetch_path(const uint8_t* key, uint32_t reads) const {
        uint32_t cur = 0;
        for (uint32_t depth = 0; depth < key_len; ++depth) {
            const vk::HNode& n = nodes[cur];
            if (reads-- == 0) {  // next read: this node's record (112 bytes: two or three lines)
                __builtin_prefetch(&n);
                __builtin_prefetch(reinterpret_cast<const char*>(&n) + 64);
                __builtin_prefetch(reinterpret_cast<const char*>(&n) + sizeof(vk::HNode) - 1);
                return;
            }
            if (!n.internal) {
                __builtin_prefetch(stems.data() + n.stem_off);
                return;
            }
            int32_t c = -1;
            if (n.table_id >= 0) {
                if (reads-- == 0) {  // next read: the child slot of the direct table
                    __builtin_prefetch(table(n) + key[depth]);
                    return;
                }
                c = table(n)[key[depth]];
            } else {
                for (uint32_t j = 0; j < n.nk; ++j)
                    if (n.unit[j] == key[depth]) c = (int32_t)n.kid[j];
            }
            if (c < 0) return;
            cur = (uint32_t)c;
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

// f(begin, end, part) over `parts` contiguous ranges of [0, n) on host threads (part 0 on the caller's)
template <class F>
static void parallel_ranges(size_t n, unsigned parts, F&& f, size_t min_n = 4096) {
    if (parts <= 1 || n < min_n) {
        f((size_t)0, n, 0u);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(parts - 1);
    for (unsigned p = 1; p < parts; ++p) th.emplace_back([&, p] { f(n * p / parts, n * (p + 1) / parts, p); });
    f((size_t)0, n / parts, 0u);
    for (auto& x : th) x.join();
}
static unsigned host_parts() {
    static unsigned parts = 0;
    if (!parts) {
        const char* e = getenv("VKZG_TREE_THREADS");
        unsigned hw = std::thread::hardware_concurrency();
        parts = e ? (unsigned)atoi(e) : (hw > 8 ? 8u : (hw ? hw : 1u));
        if (!parts) parts = 1;
    }
    return parts;
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
    void term_child(uint16_t s, int32_t c) {  // internal nodes: no literal travels
        slot.push_back(s);
        child.push_back(c);
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

// the sequential insertion loop (with the staged software prefetch); returns the number of pairs inserted
static uint64_t insert_run(vkzg_tree* t, const uint8_t* keys, const uint8_t* values, const uint32_t* order, uint64_t n) {
    const uint64_t kl = t->key_len;
    auto at = [&](uint64_t i) { return order ? (uint64_t)order[i] : i; };
    for (uint64_t i = 0; i < n; ++i) {
        // staged look-ahead: deeper levels of nearer keys (each stage reads only what an earlier stage prefetched)
        // (root and the level-1 nodes stay cached: reads 0..3 are hits; 4 = level-2 header, 5 = its slot, 6 = the
        // level-3 node (usually the extension), whose stem is the last line an insertion compares)
        if (i + 16 < n) t->prefetch_path(keys + at(i + 16) * kl, 4);
        if (i + 12 < n) t->prefetch_path(keys + at(i + 12) * kl, 5);
        if (i + 8 < n) t->prefetch_path(keys + at(i + 8) * kl, 6);
        if (i + 4 < n) t->prefetch_path(keys + at(i + 4) * kl, 7);
        if (!t->insert(keys + at(i) * kl, values + at(i) * 32)) return i;
    }
    return n;
}

// Bulk load into an EMPTY tree on host threads.  The subtrees under different first units never interact (Node::insert walks
// down from the root by key[0]), so the pairs are grouped by key[0] (stable: the order inside a group is the caller's), every
// thread builds the subtrees of a contiguous range of first units in a tree of its own with the very same insert code, and
// the partial trees are spliced under the real root (node ids, stem offsets and table ids re-based).  Any pair the
// reference would panic on makes the whole load fall back to the sequential loop, which reports the exact count.
static bool insert_parallel(vkzg_tree* t, const uint8_t* keys, const uint8_t* values, uint64_t n, unsigned parts) {
    const uint64_t kl = t->key_len;
    const bool timing = getenv("VKZG_TREE_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto t0 = now();
    std::vector<uint64_t> cnt(257, 0);
    for (uint64_t i = 0; i < n; ++i) ++cnt[keys[i * kl] + 1];
    for (int u = 0; u < 256; ++u) cnt[u + 1] += cnt[u];
    std::vector<uint32_t> order(n);
    {
        std::vector<uint64_t> cur(cnt.begin(), cnt.end() - 1);
        for (uint64_t i = 0; i < n; ++i) order[cur[keys[i * kl]]++] = (uint32_t)i;
    }
    // contiguous ranges of first units with about n / parts pairs each
    std::vector<uint32_t> cut(parts + 1, 256);
    cut[0] = 0;
    for (unsigned p = 1; p < parts; ++p) {
        uint32_t u = cut[p - 1];
        while (u < 256 && cnt[u] < n * p / parts) ++u;
        cut[p] = u;
    }
    auto t1 = now();
    std::vector<vkzg_tree> part(parts);
    std::vector<uint8_t> ok(parts, 1);
    auto build = [&](unsigned p) {
        vkzg_tree& lt = part[p];
        lt.key_len = t->key_len;
        lt.ext_width = t->ext_width;
        lt.nodes.emplace_back();
        const uint64_t b = cnt[cut[p]], e = cnt[cut[p + 1]], m = e - b;
        lt.nodes.reserve(m + m / 4 + 16);
        lt.stems.reserve(m * kl);
        lt.tables.reserve(256 * (m / 13 + 16));
        ok[p] = insert_run(&lt, keys, values, order.data() + b, m) == m;
    };
    {
        // (touching the final arenas on one more thread meanwhile was measured: the page faults of nine threads serialise in
        //  the kernel and the build phase goes from 25 to 82 ms — the arenas are sized after the builds instead, 19 ms)
        std::vector<std::thread> th;
        for (unsigned p = 1; p < parts; ++p) th.emplace_back(build, p);
        build(0);
        for (auto& x : th) x.join();
    }
    for (unsigned p = 0; p < parts; ++p)
        if (!ok[p]) return false;  // the tree is untouched: the sequential loop takes over
    auto t2 = now();
    // ---- splice: node 0 of every partial tree is its private root; its other nodes move to id base[p] + (local id - 1)
    std::vector<uint64_t> nbase(parts + 1), sbase(parts + 1), tbase(parts + 1);
    nbase[0] = 1;  // (the tree was empty: only its root)
    sbase[0] = 0;
    tbase[0] = 0;
    for (unsigned p = 0; p < parts; ++p) {
        nbase[p + 1] = nbase[p] + part[p].nodes.size() - 1;
        sbase[p + 1] = sbase[p] + part[p].stems.size();
        tbase[p + 1] = tbase[p] + part[p].tables.size() / 256;
    }
    if (nbase[parts] >= (1ull << 31)) return false;
    t->nodes.resize(nbase[parts]);
    t->stems.resize(sbase[parts]);
    t->tables.reserve(tbase[parts] * 256 + 256);  // (+ the root's own table, allocated by the set_child calls below: no regrowth)
    t->tables.resize(tbase[parts] * 256);
    auto t3 = now();
    auto splice = [&](unsigned p) {
        const vkzg_tree& lt = part[p];
        const int64_t shift = (int64_t)nbase[p] - 1;
        for (size_t i = 1; i < lt.nodes.size(); ++i) {
            HNode nd = lt.nodes[i];
            if (nd.internal) {
                for (uint32_t j = 0; j < nd.nk; ++j) nd.kid[j] = (uint32_t)(nd.kid[j] + shift);
                if (nd.table_id >= 0) nd.table_id += (int32_t)tbase[p];
            } else {
                nd.stem_off += sbase[p];
            }
            t->nodes[nbase[p] + i - 1] = nd;
        }
        if (!lt.stems.empty()) memcpy(t->stems.data() + sbase[p], lt.stems.data(), lt.stems.size());
        int32_t* dst = t->tables.data() + tbase[p] * 256;
        for (size_t i = 0; i < lt.tables.size(); ++i) dst[i] = lt.tables[i] < 0 ? -1 : (int32_t)(lt.tables[i] + shift);
    };
    {
        std::vector<std::thread> th;
        for (unsigned p = 1; p < parts; ++p) th.emplace_back(splice, p);
        splice(0);
        for (auto& x : th) x.join();
    }
    for (unsigned p = 0; p < parts; ++p) {
        const vkzg_tree& lt = part[p];
        const int64_t shift = (int64_t)nbase[p] - 1;
        // (a private root that outgrew its inline list owns table 0 of its tree; that table is spliced but unused)
        lt.for_each_child(lt.nodes[0], [&](uint8_t u, uint32_t c) { t->set_child(0, u, (uint32_t)(c + shift)); });
        t->n_keys += lt.n_keys;
        t->n_dirty += lt.n_dirty - 1;  // (every tree counted its own root)
    }
    if (timing)
        fprintf(stderr, "vkzg_tree_insert (threads): group %.1f ms, build %.1f ms, resize %.1f ms, splice %.1f ms\n", ms(t0, t1), ms(t1, t2),
                ms(t2, t3), ms(t3, now()));
    return true;
}

// n (key, value) pairs inserted IN ORDER; stops at the first pair the reference would panic on (VKZG_ERR_RANGE,
// *n_done = pairs inserted)
int32_t vkzg_tree_insert(vkzg_tree* t, const uint8_t* keys, const uint8_t* values, uint64_t n, uint64_t* n_done) {
    if (!t || (n && (!keys || !values))) return VKZG_ERR_ARG;
    const bool timing = getenv("VKZG_TREE_TIMING") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    static uint64_t par_min = 0;
    if (!par_min) {
        const char* e = getenv("VKZG_TREE_PAR_MIN");
        par_min = e ? strtoull(e, nullptr, 10) : 65536;
        if (!par_min) par_min = 1;
    }
    bool done = false;
    if (t->nodes.size() == 1 && t->nodes[0].nk == 0 && t->nodes[0].table_id < 0 && n >= par_min && n < (1ull << 32) && host_parts() > 1)
        done = insert_parallel(t, keys, values, n, host_parts());
    uint64_t inserted = n;
    if (!done) {
        t->nodes.reserve(t->nodes.size() + n + n / 4 + 16);
        t->stems.reserve(t->stems.size() + n * t->key_len);
        t->tables.reserve(t->tables.size() + 256 * (n / 13 + 16));
        inserted = insert_run(t, keys, values, nullptr, n);
    }
    if (timing)
        fprintf(stderr, "vkzg_tree_insert: %llu keys in %.1f ms%s\n", (unsigned long long)inserted,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(),
                done ? " (bulk load on host threads)" : "");
    if (n_done) *n_done = inserted;
    return inserted == n ? VKZG_OK : VKZG_ERR_RANGE;
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

// Node::path_to_stem (node.rs:101-119): push (prefix, unit, node) for every internal node that has a child under
// stem[depth]; an extension ends the walk with Ok, a missing child is VerkleError::InvalidPath.
int32_t vkzg_tree_path_to_stem(const vkzg_tree* t, const uint8_t* stem, uint32_t* path_len, uint32_t* node_ids, uint8_t* units,
                               vkzg_g1_affine* commitments, uint8_t* clean) {
    if (!t || !stem || !path_len || !node_ids || !units) return VKZG_ERR_ARG;
    uint32_t cur = 0, depth = 0;
    *path_len = 0;
    while (t->nodes[cur].internal) {
        if (depth >= t->key_len) return VKZG_ERR_RANGE;  // (the reference indexes the stem out of bounds here)
        int32_t c = t->child(cur, stem[depth]);
        if (c < 0) return VKZG_ERR_RANGE;
        node_ids[depth] = cur;
        units[depth] = stem[depth];
        const bool cl = t->is_clean(cur);
        if (clean) clean[depth] = cl ? 1 : 0;
        if (commitments) {
            if (cl)
                memcpy(&commitments[depth], &t->commits[cur], sizeof(affine_t));
            else
                memset(&commitments[depth], 0, sizeof(affine_t));
        }
        *path_len = ++depth;
        cur = (uint32_t)c;
    }
    return VKZG_OK;
}

// VerkleTree::commitment (lib.rs:127-129): recommit every dirty node, leaves first, return the root commitment.
// *n_committed (optional) = number of node commitments computed by this call (C1 / C2 helper vectors included).
int32_t vkzg_tree_commit(vkzg_ctx* ctx, uint32_t key_id, vkzg_tree* t, vkzg_g1_affine* root_out, uint64_t* n_committed) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW || !t || !root_out) return VKZG_ERR_ARG;
    if (k->n < 256 || k->n < t->ext_width || k->n < 4) return VKZG_ERR_RANGE;
    if (n_committed) *n_committed = 0;
    if (t->is_clean(0)) {
        memcpy(root_out, &t->commits[0], sizeof(affine_t));
        return VKZG_OK;
    }
    t->commits.resize(t->nodes.size());
    t->clean.resize(t->nodes.size(), 0);
    const bool timing = getenv("VKZG_TREE_TIMING") != nullptr;  // phase times of this call on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_start = now();
    const uint32_t W = t->ext_width;
    // literals travel as raw little-endian integers when they fit 32 bytes (the device reduces and converts them);
    // longer stems are reduced on the host and everything travels in Montgomery form
    const bool raw = t->key_len <= 32;
    auto literal = [](const uint8_t* b, size_t len) { return fr_from_le_bytes(b, len); };
    // ---- post-order over the dirty part: heights -> levels.  Level 0 = C1 / C2 helper vectors, 1 = extensions,
    //      >= 2 internal nodes by height.  Clean children enter as known commitments (prefix of the node array).
    // Short keys (raw): the dirty extensions travel in compact form (stem, leaf unit, value: 65 bytes each) and the device
    // writes the CSR rows of levels 0 and 1 itself (k_tree_ext_expand); levels[0] / levels[1] then only carry the owners.
    // The records are written straight into the context's pinned staging area: [stems | values | units], capacity = the
    // number of dirty nodes (an upper bound on the dirty extensions).
    const size_t ext_cap = raw ? (size_t)std::min<uint64_t>(t->n_dirty, t->nodes.size()) : 0;
    uint8_t* stage = nullptr;
    if (ext_cap) {
        stage = (uint8_t*)ctx->pinned_stage(0, ext_cap * 65);
        if (!stage) return VKZG_ERR_OOM;
    }
    fp_t* ext_stem = (fp_t*)stage;
    uint8_t* ext_val = stage + ext_cap * 32;
    uint8_t* ext_unit = stage + ext_cap * 64;
    size_t n_ext = 0;
    std::vector<LevelBuf> levels(2);
    if (raw) {
        levels[0].owner.reserve(ext_cap);
        levels[1].owner.reserve(ext_cap);
    } else if (t->n_commits == 0) {  // first (bulk) commit: every node is dirty, size the leaf-side levels once
        const size_t nn = t->nodes.size();
        levels[0].slot.reserve(2 * nn); levels[0].child.reserve(2 * nn); levels[0].lit.reserve(2 * nn);
        levels[0].row_ptr.reserve(nn + 1); levels[0].owner.reserve(nn);
        levels[1].slot.reserve(3 * nn); levels[1].child.reserve(3 * nn); levels[1].lit.reserve(3 * nn);
        levels[1].row_ptr.reserve(nn + 1); levels[1].owner.reserve(nn);
    }
    ++t->n_commits;
    std::vector<affine_t> known;            // commitments of clean children referenced by dirty parents
    std::vector<std::pair<uint32_t, int32_t>> handle(t->nodes.size(), {0xffffffffu, -1});  // node -> (level, row) ; level 0xfffffffe = known
    const fp_t one = fp_one<S>(), zero = fp_zero<S>();
    bool overflow = false;
    auto push_ext_compact = [&](uint32_t id, const HNode& n) {
        if (n_ext >= ext_cap) {  // cannot happen while n_dirty is maintained; never write past the staging area
            overflow = true;
            return;
        }
        fp_t st = fp_zero<S>();
        memcpy(st.l, t->stem(n), t->key_len);
        ext_stem[n_ext] = st;
        memcpy(ext_val + n_ext * 32, n.leaf_val.data(), 32);
        ext_unit[n_ext] = n.leaf_unit;
        ++n_ext;
        levels[0].owner.push_back(0xffffffffu);
        levels[1].owner.push_back(id);
        handle[id] = {1u, (int32_t)levels[1].owner.size() - 1};
    };
    // Bulk mode (a large part of the tree is dirty, e.g. the first commit after loading): ONE sequential pass over the node
    // array emits the dirty extensions in id order, so the depth-first walk below only visits the (few) internal nodes
    // and never touches an extension's memory.  With few dirty nodes the walk alone does
    // everything and the cost stays proportional to the dirty paths.
    const bool bulk = raw && (ctx->tree_flatten == 1 || (ctx->tree_flatten == 0 && t->n_dirty * 8 > t->nodes.size()));
    if (bulk) {
        // two passes on host threads over contiguous id ranges: count the dirty extensions of a range, then every range
        // writes its records at its prefix offset — the same id order (hence the same rows) as one sequential pass
        const size_t nn = t->nodes.size();
        const unsigned parts = host_parts();
        std::vector<size_t> cnt(parts + 1, 0);
        parallel_ranges(nn, parts, [&](size_t b, size_t e, unsigned p) {
            size_t c = 0;
            for (size_t id = b; id < e; ++id)
                if (!t->clean[id] && !t->nodes[id].internal) ++c;  // a clean node's record is not even read
            cnt[p + 1] = c;
        });
        for (unsigned p = 0; p < parts; ++p) cnt[p + 1] += cnt[p];
        n_ext = cnt[parts];
        if (n_ext > ext_cap) return VKZG_ERR_RANGE;  // cannot happen while n_dirty is maintained; never write past the staging area
        levels[0].owner.assign(n_ext, 0xffffffffu);
        levels[1].owner.resize(n_ext);
        const uint32_t kl = t->key_len;
        parallel_ranges(nn, parts, [&](size_t b, size_t e, unsigned p) {
            size_t j = cnt[p];
            for (size_t id = b; id < e; ++id) {
                if (t->clean[id]) continue;
                const HNode& n = t->nodes[id];
                if (n.internal) continue;
                fp_t st = fp_zero<S>();
                memcpy(st.l, t->stem(n), kl);
                ext_stem[j] = st;
                memcpy(ext_val + j * 32, n.leaf_val.data(), 32);
                ext_unit[j] = n.leaf_unit;
                levels[1].owner[j] = (uint32_t)id;
                handle[id] = {1u, (int32_t)j};
                ++j;
            }
        });
    }
    auto t_bulk = now();
    // depth-first walk from `start`: rows go to `levels`, commitments of clean children to `known` (+ their node ids to
    // `known_ids` when the caller has to re-base them after a merge)
    auto walk = [&](uint32_t start, std::vector<LevelBuf>& levels, std::vector<affine_t>& known, std::vector<uint32_t>* known_ids) {
    std::vector<std::pair<uint32_t, bool>> stack{{start, false}};
    while (!stack.empty()) {
        auto [id, done] = stack.back();
        stack.pop_back();
        if (t->clean[id]) {
            known.push_back(t->commits[id]);
            handle[id] = {0xfffffffeu, (int32_t)known.size() - 1};
            if (known_ids) known_ids->push_back(id);
            continue;
        }
        HNode& n = t->nodes[id];
        if (!n.internal && raw) {
            if (!bulk) push_ext_compact(id, n);
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
        // one scan of the children; a second visit (after the children are placed) only if some child still has to be
        // walked — nodes whose children are all extensions or cached are finished on the first visit
        uint32_t level = 2, nc = 0, pending = 0;
        uint8_t cu[256];
        uint32_t cc[256];
        t->for_each_child(n, [&](uint8_t u, uint32_t c) {
            cu[nc] = u;
            cc[nc++] = c;
        });
        // the children's handles are scattered over the node-id space (ids follow insertion order): request them all before
        // the first one is read, so that the misses overlap instead of queueing (63 -> see DESIGN.md ms at 2^20 keys)
        for (uint32_t j = 0; j < nc; ++j) __builtin_prefetch(&handle[cc[j]]);
        if (!done) {
            for (uint32_t j = 0; j < nc; ++j)
                if (handle[cc[j]].first == 0xffffffffu) {  // not placed yet (bulk: the dirty extensions already have their row)
                    if (!pending++) stack.push_back({id, true});
                    stack.push_back({cc[j], false});
                }
            if (pending) continue;
        }
        for (uint32_t j = 0; j < nc; ++j) {
            auto h = handle[cc[j]];
            if (h.first != 0xfffffffeu) level = std::max(level, h.first + 1);
        }
        if (levels.size() <= level) levels.resize(level + 1);
        // child reference encoded as (level << 40 | row) is too wide for int32: store node id, resolve below
        for (uint32_t j = 0; j < nc; ++j) levels[level].term_child(cu[j], (int32_t)cc[j]);
        handle[id] = {level, (int32_t)levels[level].close(id)};
    }
    };
    // Bulk mode with a wide root (the first commit of a loaded tree): the subtrees under the root's children are walked on
    // host threads into thread-local rows and merged (rows re-based, handles shifted); the extensions already have their
    // rows from the bulk pass, so the walks share nothing but the handle array, which they write at disjoint node ids.
    // A walk is ~1 us of dependent cache misses per internal node (65 k of them at 2^20 keys: 63 ms on one thread).
    {
        std::vector<uint32_t> subs;
        if (bulk && host_parts() > 1 && !t->clean[0] && t->nodes[0].internal)
            t->for_each_child(t->nodes[0], [&](uint8_t, uint32_t c) {
                if (!t->clean[c] && t->nodes[c].internal) subs.push_back(c);
            });
        const unsigned parts = (unsigned)std::min<size_t>(host_parts(), subs.size() / 4);
        if (parts > 1) {
            std::vector<std::vector<LevelBuf>> lv(parts);
            std::vector<std::vector<affine_t>> kn(parts);
            std::vector<std::vector<uint32_t>> kid(parts);
            parallel_ranges(subs.size(), parts, [&](size_t b, size_t e, unsigned p) {
                for (size_t i = b; i < e; ++i) walk(subs[i], lv[p], kn[p], &kid[p]);
            }, /*min_n=*/0);
            for (unsigned p = 0; p < parts; ++p) {
                const int32_t koff = (int32_t)known.size();
                known.insert(known.end(), kn[p].begin(), kn[p].end());
                if (koff)
                    for (uint32_t id : kid[p]) handle[id].second += koff;
                if (levels.size() < lv[p].size()) levels.resize(lv[p].size());
                for (size_t l = 2; l < lv[p].size(); ++l) {
                    LevelBuf& D = levels[l];
                    const LevelBuf& Sx = lv[p][l];
                    const int32_t roff = (int32_t)D.owner.size();
                    const uint32_t toff = (uint32_t)D.slot.size();
                    D.slot.insert(D.slot.end(), Sx.slot.begin(), Sx.slot.end());
                    D.child.insert(D.child.end(), Sx.child.begin(), Sx.child.end());
                    for (size_t r = 1; r < Sx.row_ptr.size(); ++r) D.row_ptr.push_back(Sx.row_ptr[r] + toff);
                    D.owner.insert(D.owner.end(), Sx.owner.begin(), Sx.owner.end());
                    if (roff)
                        for (uint32_t id : Sx.owner) handle[id].second += roff;
                }
            }
        }
    }
    walk(0u, levels, known, nullptr);
    if (overflow) return VKZG_ERR_RANGE;
    auto t_flat = now();
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
        for (size_t j = 0; j < nn; ++j) {
            uint32_t r = perm[j];
            for (uint32_t t2 = L.row_ptr[r]; t2 < L.row_ptr[r + 1]; ++t2) S.term_child(L.slot[t2], L.child[t2]);
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
    for (size_t l = 2; l < levels.size(); ++l) {
        std::vector<int32_t>& ch = levels[l].child;
        parallel_ranges(ch.size(), host_parts(), [&](size_t b, size_t e, unsigned) {  // (scattered reads of the handle array)
            for (size_t i = b; i < e; ++i) {
                if (i + 16 < e) __builtin_prefetch(&handle[(uint32_t)ch[i + 16]]);
                auto h = handle[(uint32_t)ch[i]];
                ch[i] = h.first == 0xfffffffeu ? h.second : (int32_t)(base[h.first] + (uint64_t)h.second);
            }
        });
    }
    auto t_ids = now();
    // ---- device passes
    DevBuf<affine_t> all;
    VK_TRY(all.alloc(ctx, total));
    if (!known.empty()) VK_CUDA(cudaMemcpyAsync(all.p, known.data(), known.size() * sizeof(affine_t), cudaMemcpyHostToDevice, ctx->stream));
    if (n_ext) {
        DevBuf<fp_t> d_st;
        DevBuf<uint8_t> d_un, d_va;
        VK_TRY(upload(ctx, d_st, ext_stem, n_ext));
        VK_TRY(upload(ctx, d_un, ext_unit, n_ext));
        VK_TRY(upload(ctx, d_va, ext_val, n_ext * 32));
        VK_TRY(tree_ext_levels(ctx, *k, d_st, d_un, d_va, n_ext, W, all.p, base[0], base[1]));
        VK_TRY(stream_sync(ctx));
    }
    for (size_t l = raw ? 2 : 0; l < levels.size(); ++l) {
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
        if (!L.lit.empty()) VK_TRY(upload(ctx, d_li, L.lit.data(), L.lit.size()));  // internal levels carry no literals
        VK_TRY(tree_level(ctx, *k, d_rp, nn, d_sl, d_ch, d_li, L.slot.size(), all.p, all.p + base[l], raw));
        VK_TRY(stream_sync(ctx));  // the host vectors of this level go out of use before the next upload reuses the pool
    }
    auto t_dev = now();
    // ---- cache the new commitments on the host
    // (the level-0 helper rows C1 / C2 belong to no node and stay on the device)
    const size_t n_new = total - base[1];
    const affine_t* host = (const affine_t*)ctx->pinned_stage(1, n_new * sizeof(affine_t));
    if (!host) return VKZG_ERR_OOM;
    VK_CUDA(cudaMemcpyAsync((void*)host, all.p + base[1], n_new * sizeof(affine_t), cudaMemcpyDeviceToHost, ctx->stream));
    VK_TRY(stream_sync(ctx));
    uint64_t off = 0;
    for (size_t l = 1; l < levels.size(); ++l) {
        const std::vector<uint32_t>& own = levels[l].owner;  // node ids are distinct: the ranges write disjoint entries
        parallel_ranges(own.size(), host_parts(), [&](size_t b, size_t e, unsigned) {
            for (size_t j = b; j < e; ++j) {
                uint32_t id = own[j];
                if (id != 0xffffffffu) {
                    t->commits[id] = host[off + j];
                    t->clean[id] = 1;
                }
            }
        });
        off += own.size();
    }
    if (timing) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "vkzg_tree_commit: %zu rows; flatten %.1f ms (setup + bulk pass %.1f, walk %.1f), sort+ids %.1f ms, upload+kernels %.1f ms, cache-back %.1f ms\n",
                n_new, ms(t_start, t_flat), ms(t_start, t_bulk), ms(t_bulk, t_flat), ms(t_flat, t_ids), ms(t_ids, t_dev), ms(t_dev, now()));
    }
    t->n_dirty = 0;
    if (n_committed) *n_committed = n_new;
    memcpy(root_out, &t->commits[0], sizeof(affine_t));
    return VKZG_OK;
}

}  // extern "C"
