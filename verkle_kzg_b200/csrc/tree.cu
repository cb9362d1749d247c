// T1: Node::gen_commitment (verkle-tree/src/node.rs:212-277), level-synchronous.
//
// The reference walks the tree depth first and commits every node with a dense width-256 vector that is
// almost entirely zero.  A parent's scalars are to_data_item(child commitment) (lib.rs:56-67), so nodes of
// one depth are independent: the tree is committed one LEVEL per launch pair, leaves first, each node a
// sparse term list (slot, scalar):
//   k_tree_scalars    per term: the literal scalar, or to_data_item of the child's commitment from the
//                     previous level (affine -> compressed bytes -> mod r)
//   k_fixed_base_msm  (commit.cu, CSR mode) one warp per node over the key's window tables
//   k_normalize       batched affine normalisation (shared inversions)
// The host flattens the pointer tree into these per-level CSR arrays (verkle_kzg_b200/tree.py mirrors
// Node::insert); the HashMap iteration-order dependence of the reference's extension layout (quirk Q6)
// never reaches the device, which only sees explicit slots.
#include "vk_common.cuh"

namespace vk {

// lit_raw: literals are 32-byte little-endian INTEGERS (from_le_bytes_mod_order input, any value < 2^256) instead of
// Montgomery field elements; one product by R^2 reduces and converts them here instead of on the host
__global__ void __launch_bounds__(128) k_tree_scalars(const int32_t* __restrict__ child, const fp_t* __restrict__ lit,
                                                      const affine_t* __restrict__ prev, uint64_t n_terms, bool lit_raw,
                                                      fp_t* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    int32_t c = child[t];
    fp_t r;
    if (c < 0) {
        r = fp_load_ro(lit + t);
        if (lit_raw) {
            fp_t r2;
#pragma unroll
            for (int k = 0; k < 8; ++k) r2.l[k] = S::r2(k);
            r = fp_mul<S>(r2, r);
        }
    } else {
        affine_t p;
        p.x = fp_load(&prev[c].x);
        p.y = fp_load(&prev[c].y);
        r = fp_zero<S>();
        if (!affine_is_inf(p)) {
            fp_t v, r2;
            affine_compress(p, v.l);
#pragma unroll
            for (int k = 0; k < 8; ++k) r2.l[k] = S::r2(k);
            r = fp_mul<S>(r2, v);
        }
    }
    fp_store(out + t, r);
}

int32_t tree_level(vkzg_ctx* ctx, const Key& k, const uint32_t* d_row_ptr, uint64_t n_nodes, const uint16_t* d_slot,
                   const int32_t* d_child, const fp_t* d_lit, uint64_t n_terms, const affine_t* d_prev, affine_t* d_out, bool lit_raw) {
    if (n_nodes == 0) return VKZG_OK;
    DevBuf<fp_t> sc;
    DevBuf<xyzz_t> acc;
    VK_TRY(sc.alloc(ctx, n_terms));
    VK_TRY(acc.alloc(ctx, n_nodes));
    if (n_terms) {
        k_tree_scalars<<<ceil_div_u64(n_terms, 128), 128, 0, ctx->stream>>>(d_child, d_lit, d_prev, n_terms, lit_raw, sc);
        VK_TRY(launch_check(ctx));
    }
    // lanes per node by the level's mean number of terms (leaf-side levels have 2-3 terms, internal nodes up to 256)
    uint64_t avg_terms = n_terms / n_nodes;
    static int lpj_small = -1;
    if (lpj_small < 0) {
        const char* e = getenv("VKZG_TREE_LPJ");  // tuning knob for the 2-4-term levels (1, 2 or 4 lanes per node)
        lpj_small = e ? atoi(e) : 1;
    }
    uint32_t lpj = avg_terms <= 4 ? (uint32_t)lpj_small : (avg_terms <= 48 ? 8 : 32);
    if (n_nodes < 65536 && lpj < 4) lpj = 4;  // few nodes: parallelism over lanes matters more than the fold
    // a handful of wide nodes (the top of the tree: up to 256 children each): slice every node over several warps
    uint32_t split = 1;
    if (avg_terms >= 32 && n_nodes * 4 <= (uint64_t)ctx->sm_count * 16) {
        split = (uint32_t)(avg_terms / 8);
        uint64_t room = (uint64_t)ctx->sm_count * 16 / n_nodes;
        if (split > room) split = (uint32_t)room;
        if (split > 32) split = 32;
        if (split < 1) split = 1;
    }
    VK_TRY(fixed_base_msm_csr(ctx, k, sc, 0, n_nodes, 0, 0xffffffffu, d_row_ptr, d_slot, acc, lpj, split));
    return normalize_points(ctx, acc, n_nodes, d_out);
}

// Extension nodes in compact form (one stem, one leaf unit, one 32-byte value each; the stem is the whole key, so every
// extension of the reference holds exactly one leaf, node.rs:173-177) -> the CSR rows of the two leaf-side levels:
//   level 0, row j: the C1 / C2 helper vector (node.rs:226-240): value halves at slots (2 idx) % W and (2 idx + 1) % W
//   level 1, row j: commit([1, stem, C1, C2]) (node.rs:243-253) with only the helper that exists (C1 iff idx < W/2)
// Literals are raw little-endian integers (k_tree_scalars converts them).  T0 = terms per level-0 row (1 iff W == 1:
// both halves land on slot 0 and the later write, the high half, wins).
__global__ void __launch_bounds__(128) k_tree_ext_expand(const fp_t* __restrict__ stem, const uint8_t* __restrict__ unit,
                                                         const uint4* __restrict__ val, uint64_t n, uint32_t W, int32_t base0,
                                                         uint32_t* __restrict__ rp0, uint16_t* __restrict__ slot0,
                                                         int32_t* __restrict__ child0, fp_t* __restrict__ lit0,
                                                         uint32_t* __restrict__ rp1, uint16_t* __restrict__ slot1,
                                                         int32_t* __restrict__ child1, fp_t* __restrict__ lit1) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n) return;
    const uint32_t T0 = W == 1 ? 1u : 2u;
    rp0[j] = (uint32_t)(T0 * j);
    rp1[j] = (uint32_t)(3 * j);
    if (j == n) return;
    uint32_t idx = unit[j];
    uint4 lo = val[2 * j], hi = val[2 * j + 1];
    fp_t v = fp_zero<S>();
    uint64_t t = T0 * j;
    if (T0 == 2) {
        v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
        slot0[t] = (uint16_t)((2 * idx) % W);
        child0[t] = -1;
        fp_store(lit0 + t, v);
        ++t;
    }
    v.l[0] = hi.x; v.l[1] = hi.y; v.l[2] = hi.z; v.l[3] = hi.w;
    slot0[t] = (uint16_t)((2 * idx + 1) % W);
    child0[t] = -1;
    fp_store(lit0 + t, v);
    t = 3 * j;
    v = fp_zero<S>();
    v.l[0] = 1;
    slot1[t] = 0; child1[t] = -1; fp_store(lit1 + t, v);
    slot1[t + 1] = 1; child1[t + 1] = -1; fp_store(lit1 + t + 1, fp_load_ro(stem + j));
    slot1[t + 2] = idx < W / 2 ? 2 : 3; child1[t + 2] = base0 + (int32_t)j;  // lit1 unused for child terms
}

// all[base0 .. base0+n) <- helper commitments, all[base1 .. base1+n) <- extension commitments
int32_t tree_ext_levels(vkzg_ctx* ctx, const Key& k, const fp_t* d_stem, const uint8_t* d_unit, const uint8_t* d_val, uint64_t n,
                        uint32_t W, affine_t* d_all, uint64_t base0, uint64_t base1) {
    if (n == 0) return VKZG_OK;
    const uint64_t T0 = W == 1 ? 1 : 2;
    DevBuf<uint32_t> rp0, rp1;
    DevBuf<uint16_t> sl0, sl1;
    DevBuf<int32_t> ch0, ch1;
    DevBuf<fp_t> li0, li1;
    VK_TRY(rp0.alloc(ctx, n + 1));
    VK_TRY(rp1.alloc(ctx, n + 1));
    VK_TRY(sl0.alloc(ctx, T0 * n));
    VK_TRY(ch0.alloc(ctx, T0 * n));
    VK_TRY(li0.alloc(ctx, T0 * n));
    VK_TRY(sl1.alloc(ctx, 3 * n));
    VK_TRY(ch1.alloc(ctx, 3 * n));
    VK_TRY(li1.alloc(ctx, 3 * n));
    k_tree_ext_expand<<<ceil_div_u64(n + 1, 128), 128, 0, ctx->stream>>>(d_stem, d_unit, (const uint4*)d_val, n, W, (int32_t)base0, rp0, sl0,
                                                                         ch0, li0, rp1, sl1, ch1, li1);
    VK_TRY(launch_check(ctx));
    VK_TRY(tree_level(ctx, k, rp0, n, sl0, ch0, li0, T0 * n, d_all, d_all + base0, true));
    return tree_level(ctx, k, rp1, n, sl1, ch1, li1, 3 * n, d_all, d_all + base1, true);
}

}  // namespace vk

using namespace vk;

extern "C" {

int32_t vkzg_tree_level_dev(vkzg_ctx* ctx, uint32_t key_id, const uint32_t* d_row_ptr, uint64_t n_nodes, const uint16_t* d_slot,
                            const int32_t* d_child, const vkzg_fr* d_lit, uint64_t n_terms, const vkzg_g1_affine* d_prev,
                            vkzg_g1_affine* d_out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (n_nodes && (!d_row_ptr || !d_out)) return VKZG_ERR_ARG;
    if (n_terms && (!d_slot || !d_child || !d_lit)) return VKZG_ERR_ARG;
    return tree_level(ctx, *k, d_row_ptr, n_nodes, d_slot, d_child, (const fp_t*)d_lit, n_terms, (const affine_t*)d_prev,
                      (affine_t*)d_out, false);
}

int32_t vkzg_tree_commit_levels(vkzg_ctx* ctx, uint32_t key_id, uint32_t n_levels, const uint64_t* nodes_per_level,
                                const uint32_t* const* row_ptr, const uint16_t* const* slot, const int32_t* const* child,
                                const vkzg_fr* const* lit, vkzg_g1_affine* root_out) {
    VK_TRY(ctx_check(ctx));
    Key* k = ctx->key(key_id);
    if (!k || k->kind != VKZG_KEY_WINDOW) return VKZG_ERR_ARG;
    if (!n_levels || !nodes_per_level || !row_ptr || !slot || !child || !lit || !root_out) return VKZG_ERR_ARG;
    if (nodes_per_level[n_levels - 1] != 1) return VKZG_ERR_ARG;  // the last level is the root
    // validate on the host: slots inside the key, children among the nodes of earlier levels
    uint64_t total_nodes = 0;
    for (uint32_t l = 0; l < n_levels; ++l) {
        uint64_t nn = nodes_per_level[l];
        if (!row_ptr[l]) return VKZG_ERR_ARG;
        uint64_t nt = row_ptr[l][nn];
        for (uint64_t t = 0; t < nt; ++t) {
            if (slot[l][t] >= k->n) return VKZG_ERR_RANGE;
            if (child[l][t] >= 0 && (uint64_t)child[l][t] >= total_nodes) return VKZG_ERR_RANGE;
        }
        total_nodes += nn;
    }
    if (total_nodes >= (1ull << 31)) return VKZG_ERR_RANGE;
    DevBuf<affine_t> all;
    VK_TRY(all.alloc(ctx, total_nodes));
    uint64_t off = 0;
    for (uint32_t l = 0; l < n_levels; ++l) {
        uint64_t nn = nodes_per_level[l];
        uint64_t nt = row_ptr[l][nn];
        DevBuf<uint32_t> d_rp;
        DevBuf<uint16_t> d_sl;
        DevBuf<int32_t> d_ch;
        DevBuf<fp_t> d_li;
        VK_TRY(upload(ctx, d_rp, row_ptr[l], nn + 1));
        VK_TRY(upload(ctx, d_sl, slot[l], nt));
        VK_TRY(upload(ctx, d_ch, child[l], nt));
        VK_TRY(upload(ctx, d_li, lit[l], nt));
        VK_TRY(tree_level(ctx, *k, d_rp, nn, d_sl, d_ch, d_li, nt, all.p, all.p + off, false));
        off += nn;
    }
    const affine_t* a = all.p + (total_nodes - 1);
    VK_TRY(download(ctx, root_out, a, 1));
    return stream_sync(ctx);
}

}  // extern "C"
