/*
 * vkzg.h — C ABI of libvkzg.so: the B200 (sm_100a) implementation of the vector-commitment hot path of
 * SleepingShell/verkle-kzg.  The reference has no FFI; its operator interface for this path is the pair
 * of Rust traits VectorCommitment (vector-commit/src/lib.rs:70-174) and VectorCommitmentMultiproof
 * (vector-commit/src/multiproof.rs:90-216).  Each entry point below names the trait method / function
 * whose body it replaces.  INTEGRATION.md shows the Rust `extern "C"` block and the trait impls that
 * bind them.
 *
 * Conventions
 *   - every function returns a status: 0 = ok, < 0 = error (vkzg_strerror); nothing throws across
 *     the boundary; the caller owns every buffer; host-pointer calls return after the result is in
 *     the caller's buffer; `_dev` calls take DEVICE pointers, enqueue on the context's stream and
 *     return without synchronising.
 *   - field elements (vkzg_fr / vkzg_fq): 8 little-endian u32 limbs holding the MONTGOMERY form with
 *     R = 2^256 — byte-identical to ark-ff 0.4 `Fp<MontBackend<_,4>>` (4 x u64), so a
 *     `&[Fr]` can be passed as `*const vkzg_fr` with no conversion.
 *   - points (vkzg_g1_affine): x || y, 64 bytes, (0,0) encodes the point at infinity (ark-ec
 *     `Affine::identity()` has `infinity = true`; the shim maps it).  All point results are canonical
 *     affine coordinates, hence independent of any internal representation or summation order.
 *   - one vkzg_ctx per host thread and per GPU (or externally synchronised).
 *   - there is NO CPU fallback: every call fails with VKZG_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef VKZG_H
#define VKZG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vkzg_ctx vkzg_ctx;
typedef struct { uint32_t l[8]; } vkzg_fr;
typedef struct { uint32_t l[8]; } vkzg_fq;
typedef struct { vkzg_fq x, y; } vkzg_g1_affine;

enum {
    VKZG_OK = 0,
    VKZG_ERR_CUDA = -1,        /* a CUDA call failed / no device                                           */
    VKZG_ERR_ARG = -2,         /* null pointer, zero size, bad key id                                      */
    VKZG_ERR_RANGE = -3,       /* index / width outside the key (reference: panic or Err(OutOfCRS))        */
    VKZG_ERR_UNSUPPORTED = -4, /* e.g. width not a power of two for IPA (reference: unwrap on odd split)   */
    VKZG_ERR_OOM = -5
};

/* key kinds (vkzg_key_load `kind`) */
enum {
    VKZG_KEY_WINDOW = 1, /* fixed-base signed-window tables for every base: batched width-N commits, IPA, KZG opens */
    VKZG_KEY_MSM = 2     /* 2^(c*w) multiples of every base: one large shared-bucket Pippenger MSM                   */
};

const char* vkzg_strerror(int32_t status);
uint32_t vkzg_abi_version(void);

/* ---- context ---------------------------------------------------------------------------------------- */
int32_t vkzg_ctx_create(vkzg_ctx** out, int32_t device_id);
/* same, but all work is enqueued on the caller's cudaStream_t (e.g. torch's current stream) */
int32_t vkzg_ctx_create_on_stream(vkzg_ctx** out, int32_t device_id, void* cuda_stream);
int32_t vkzg_ctx_destroy(vkzg_ctx* ctx);
int32_t vkzg_ctx_sync(vkzg_ctx* ctx);
/* options: VKZG_OPT_IPA_TWO_STREAMS (default 1): batches of >= 8192 IPA proofs run as two half-batches on two streams so
 * that one half's latency-bound challenge / fold kernels hide under the other half's MSM kernel.
 * VKZG_OPT_TREE_FLATTEN (default 0 = automatic): how vkzg_tree_commit gathers the dirty nodes — 1 = always the sequential
 * bulk pass over the node array, 2 = always the depth-first walk of the dirty paths (automatic: bulk when more than an
 * eighth of the nodes is dirty).  Results are identical; the knob exists for tests and measurements.
 * VKZG_OPT_BATCH_AFFINE (default -1 = automatic): big batches of dense width-N jobs (commits, IPA cross terms) are summed by the
 * batch-affine tree (6 field products per addition, shared inversions) instead of per-lane XYZZ accumulators (10 products);
 * 0 = never, 1 = whenever the shape allows.  Results are identical (canonical affine points).
 * VKZG_OPT_MULTIPROOF_CHECK_Y (default 0; diagnostic): vkzg_multiproof_verify_ipa additionally compares the proof's
 * evaluation with g2(t) = sum_q r^q y_q / (t - z_q), the sum the reference computes and never uses (multiproof.rs:201-215:
 * its verifier accepts ANY claimed y_q).  The comparison cannot be the default: the reference's prover divides by
 * (X - w^z) in g but by (t - z), z an INTEGER, in h (multiproof.rs:155-166, quirk Q4), so (h - g)(t) != g2(t) even for an
 * honest proof — with the option on, every proof made by the reference's algorithm (and by this library, which is
 * bit-exact with it) is rejected.  See DESIGN.md section 6.                                                            */
enum { VKZG_OPT_IPA_TWO_STREAMS = 1, VKZG_OPT_TREE_FLATTEN = 2, VKZG_OPT_MULTIPROOF_CHECK_Y = 3, VKZG_OPT_BATCH_AFFINE = 4 };
int32_t vkzg_ctx_set_option(vkzg_ctx* ctx, int32_t option, int32_t value);
/* Scratch memory comes from a stream-ordered pool private to the context (the device's default pool is not touched); freed
 * blocks stay cached between calls (VKZG_POOL_KEEP_MB in the environment bounds that) — vkzg_ctx_trim synchronises and
 * returns them to the driver, vkzg_ctx_destroy does the same.                                                          */
int32_t vkzg_ctx_trim(vkzg_ctx* ctx);
/* kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t vkzg_ctx_launches(const vkzg_ctx* ctx);

/* ---- keys: KZGKey.lagrange_commitments (kzg/mod.rs:27-57) / IPAUniversalParams{g,q} (ipa/mod.rs:22-52) --- */
/* `q` may be NULL (KZG).  window_bits = 0 picks the default (16 for VKZG_KEY_WINDOW, by size for MSM).      */
int32_t vkzg_key_load(vkzg_ctx* ctx, const vkzg_g1_affine* bases, uint32_t n, const vkzg_g1_affine* q,
                      uint32_t kind, uint32_t window_bits, uint32_t* key_id);
int32_t vkzg_key_load_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_bases, uint32_t n, const vkzg_g1_affine* d_q,
                          uint32_t kind, uint32_t window_bits, uint32_t* key_id);
int32_t vkzg_key_free(vkzg_ctx* ctx, uint32_t key_id);
uint64_t vkzg_key_table_bytes(const vkzg_ctx* ctx, uint32_t key_id);

/* ---- M1: utils::inner_product G x F (utils.rs:16-19) as used by KZG::commit (kzg/mod.rs:126-134) and
 *      IPA::commit (ipa/mod.rs:130-135) ------------------------------------------------------------------- */
/* one large MSM over the first n bases of a VKZG_KEY_MSM key (n <= key size: zip truncation, quirk Q1) */
int32_t vkzg_msm(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* scalars, uint64_t n, vkzg_g1_affine* out);
int32_t vkzg_msm_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_scalars, uint64_t n, vkzg_g1_affine* d_out);
/* same over the slice [first, first + n) of the key's bases (point-range sharding across GPUs) */
int32_t vkzg_msm_range_dev(vkzg_ctx* ctx, uint32_t key_id, uint64_t first, const vkzg_fr* d_scalars, uint64_t n,
                           vkzg_g1_affine* d_out);
/* B independent commits of width w (w <= key size) — scalars[B][w] -> out[B] */
int32_t vkzg_commit_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* scalars, uint32_t w, uint64_t B,
                          vkzg_g1_affine* out);
int32_t vkzg_commit_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_scalars, uint32_t w, uint64_t B,
                              vkzg_g1_affine* d_out);
/* sum of n points (the combine step after an all-gather of per-rank partial commitments) */
int32_t vkzg_g1_sum_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_points, uint64_t n, vkzg_g1_affine* d_out);
int32_t vkzg_g1_sum(vkzg_ctx* ctx, const vkzg_g1_affine* points, uint64_t n, vkzg_g1_affine* out);

/* ---- D1: VCCommitment::to_data_item (vector-commit/src/lib.rs:56-67) -------------------------------- */
int32_t vkzg_to_data_item(vkzg_ctx* ctx, const vkzg_g1_affine* points, uint64_t n, vkzg_fr* out);
int32_t vkzg_to_data_item_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_points, uint64_t n, vkzg_fr* d_out);

/* ---- L1 / I2: element-wise Fr vector arithmetic — LagrangeBasis AddAssign / Sub / Mul<F> (lagrange_basis.rs:202-233),
 *      utils::elementwise_mul and vec_add_and_distribute (utils.rs:21-38).
 *      op: 0 a + b, 1 a - b, 2 a .* b, 3 a * x, 4 a + x * b.  `x` is a HOST pointer to one scalar (ops 3, 4).        */
int32_t vkzg_fr_vector_op(vkzg_ctx* ctx, int32_t op, const vkzg_fr* a, const vkzg_fr* b, const vkzg_fr* x, uint64_t n, vkzg_fr* out);
int32_t vkzg_fr_vector_op_dev(vkzg_ctx* ctx, int32_t op, const vkzg_fr* d_a, const vkzg_fr* d_b, const vkzg_fr* x, uint64_t n,
                              vkzg_fr* d_out);

/* ---- B1 / E1 / K1 / K2: precompute.rs:72-90, lagrange_basis.rs:63-83, :91-119, :121-142 ---------------- */
/* barycentric coefficients of B points over the key's domain (size N): out[B][N] */
int32_t vkzg_barycentric_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* points, uint64_t B, vkzg_fr* out);
/* Data rows f[B][len] hold len <= N evaluations.  domain_n = 0: LagrangeBasis::from_vec(data), whose own
 * domain has size Dn = next_pow2(len) (lagrange_basis.rs:152-155); domain_n != 0:
 * from_vec_and_domain(data, D::new(domain_n)) (lagrange_basis.rs:24-31, as the reference's KZG tests do with
 * the key's domain), Dn = next_pow2(domain_n).  The key's precompute has size N; the reference indexes
 * out of bounds unless Dn <= N (VKZG_ERR_UNSUPPORTED otherwise).                                            */
/* f(z) with the reference's 3-way branch (quirk Q2): out[B] */
int32_t vkzg_evaluate_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n,
                            const vkzg_fr* points, uint64_t B, vkzg_fr* out);
/* quotient (f - f(z)) / (X - z) in evaluation form as KZG::prove_point selects it: in-domain branch for
 * z <= N (z == N, or z >= Dn, is the reference's out-of-bounds panic -> VKZG_ERR_RANGE), outside-domain
 * branch otherwise.  out[B][Dn], y[B]                                                                        */
int32_t vkzg_quotient_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n,
                            const vkzg_fr* points, uint64_t B, vkzg_fr* out, vkzg_fr* y);

int32_t vkzg_evaluate_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_f, uint32_t len, uint32_t domain_n,
                                const vkzg_fr* d_points, uint64_t B, vkzg_fr* d_y);
int32_t vkzg_quotient_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_f, uint32_t len, uint32_t domain_n,
                                const vkzg_fr* d_points, uint64_t B, vkzg_fr* d_out, vkzg_fr* d_y);

/* ---- K3: KZG::prove_point (kzg/mod.rs:136-154), batched over B openings ------------------------------ */
int32_t vkzg_kzg_open_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f /*[B][len]*/, uint32_t len, uint32_t domain_n,
                            const vkzg_fr* points /*[B]*/, uint64_t B, vkzg_g1_affine* proof /*[B]*/, vkzg_fr* y /*[B]*/);
/* device-pointer variant: rows that hit the reference's panic get the identity as proof (no status) */
int32_t vkzg_kzg_open_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_f, uint32_t len, uint32_t domain_n,
                                const vkzg_fr* d_points, uint64_t B, vkzg_g1_affine* d_proof, vkzg_fr* d_y);
/* KZG::commit (kzg/mod.rs:126-134) + KZG::prove_point (:136-154) over the same B vectors with ONE upload of the rows (bulk
 * callers that commit and open the same data, as benches/kzg.rs does): commitments[B], proof[B], y[B]               */
int32_t vkzg_kzg_commit_open_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f, uint32_t len, uint32_t domain_n,
                                   const vkzg_fr* points, uint64_t B, vkzg_g1_affine* commitments, vkzg_g1_affine* proof,
                                   vkzg_fr* y);

/* KZG::prove_all_points (kzg/mod.rs:200-235): every vector opened at ALL Dn points of its data domain — proof[B][Dn],
 * y[B][Dn], with proof[b][i] == KZG::prove_point(f_b, i) and y[b][i] = f_b[i] (0 for len <= i < Dn).  The reference's
 * function is private dead code that stops at Feist-Khovratovich's h-vector and panics on data[i], i >= N; its contract
 * (what its unregistered test checks: every entry verifies as a single-point proof) is what this entry keeps.          */
int32_t vkzg_kzg_prove_all_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* f /*[B][len]*/, uint32_t len, uint32_t domain_n,
                                 uint64_t B, vkzg_g1_affine* proof /*[B][Dn]*/, vkzg_fr* y /*[B][Dn]*/);

/* ---- I1: IPA::prove_point + low_level_ipa (ipa/mod.rs:137-154, :268-319), batched -------------------- */
/* a[B][N], points[B], commitments[B].  `prefix` (may be NULL) is the byte state of an in-flight transcript
 * shared by all B proofs (lib.rs:127-133 / multiproof.rs:174), `dst` the transcript's domain label
 * (NULL = "ipa").  The prefix may have any length (like the reference's transcript): beyond 160 bytes its whole
 * 64-byte blocks are hashed on the host and the device continues from that SHA-256 state.
 * Outputs: L[B][log2 N], R[B][log2 N], tip[B], y[B].                                                        */
int32_t vkzg_ipa_prove_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* a, const vkzg_fr* points,
                             const vkzg_g1_affine* commitments, uint64_t B, const uint8_t* prefix, uint32_t prefix_len,
                             const char* dst, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* y);
int32_t vkzg_ipa_prove_batch_dev(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* d_a, const vkzg_fr* d_points,
                                 const vkzg_g1_affine* d_commitments, uint64_t B, const uint8_t* prefix, uint32_t prefix_len,
                                 const char* dst, vkzg_g1_affine* d_L, vkzg_g1_affine* d_R, vkzg_fr* d_tip, vkzg_fr* d_y);
/* IPA::commit followed by IPA::prove_point on the same B vectors in ONE call (fresh "ipa" transcripts): the rows cross
 * PCIe once and the commitments stay on the device between the two steps.  Outputs: commitments[B] + the proofs.    */
int32_t vkzg_ipa_commit_prove_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* a, const vkzg_fr* points, uint64_t B,
                                    vkzg_g1_affine* commitments, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* y);
/* ---- I3: IPA::verify_point + low_level_verify_ipa (ipa/mod.rs:165-181, :321-360), batched ------------ */
/* ok[B] receives 1 (valid) / 0 (invalid) */
int32_t vkzg_ipa_verify_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* points, const vkzg_g1_affine* commitments,
                              uint64_t B, const uint8_t* prefix, uint32_t prefix_len, const char* dst,
                              const vkzg_g1_affine* L, const vkzg_g1_affine* R, const vkzg_fr* tip, const vkzg_fr* y,
                              int32_t* ok);
/* ---- I4: IPA::prove_commitment / verify_commitment_proof (ipa/mod.rs:199-265) ------------------------- */
int32_t vkzg_ipa_prove_commitment_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* a, const vkzg_g1_affine* commitments,
                                        uint64_t B, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip);
/* IPA::verify_commitment_proof (ipa/mod.rs:238-265) over B (commitment, proof) pairs of the key's full width:
 * tip * <G, s> == the folded commitment; ok[B] receives 1 / 0.  The key needs no Q.                              */
int32_t vkzg_ipa_verify_commitment_batch(vkzg_ctx* ctx, uint32_t key_id, const vkzg_g1_affine* commitments, uint64_t B,
                                         const vkzg_g1_affine* L, const vkzg_g1_affine* R, const vkzg_fr* tip, int32_t* ok);

/* ---- P1: VectorCommitmentMultiproof::prove_multiproof (multiproof.rs:99-176) -------------------------- */
/* scheme: 0 = IPA (key has q), 1 = KZG.  f[m][N], C[m], z[m], y[m].
 * IPA outputs D, L[log2 N], R[log2 N], tip, yout;  KZG outputs D, L[0] = proof point, yout.             */
int32_t vkzg_multiproof_prove(vkzg_ctx* ctx, uint32_t key_id, int32_t scheme, const vkzg_fr* f, const vkzg_g1_affine* C,
                              const uint64_t* z, const vkzg_fr* y, uint64_t m, vkzg_g1_affine* D, vkzg_g1_affine* L,
                              vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* yout);
/* same with the m x N rows already resident in device memory (d_f); the query metadata (C, z, y) and all outputs
 * stay host pointers because the outer transcript is hashed on the host               */
int32_t vkzg_multiproof_prove_dev(vkzg_ctx* ctx, uint32_t key_id, int32_t scheme, const vkzg_fr* d_f, const vkzg_g1_affine* C,
                                  const uint64_t* z, const vkzg_fr* y, uint64_t m, vkzg_g1_affine* D, vkzg_g1_affine* L,
                                  vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* yout);
/* K multiproofs in one call (bulk provers: a block's worth of multiproofs, benches/ipa.rs:111-132 in a loop).  Queries are
 * concatenated: proof k owns rows / (C, z, y) entries [sum_{j<k} m_each[j], + m_each[k]).  Outputs per proof: D[K],
 * L[K][log2 N], R[K][log2 N], tip[K], yout[K] (KZG: L[K] = proof points).  The K outer transcripts are hashed by host
 * threads, D_k / E_k are K-job commits and the K inner openings run as ONE IPA batch — byte-identical to K single calls.  */
int32_t vkzg_multiproof_prove_batch(vkzg_ctx* ctx, uint32_t key_id, int32_t scheme, const vkzg_fr* f, const vkzg_g1_affine* C,
                                    const uint64_t* z, const vkzg_fr* y, const uint64_t* m_each, uint64_t K, vkzg_g1_affine* D,
                                    vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* yout);
int32_t vkzg_multiproof_prove_batch_dev(vkzg_ctx* ctx, uint32_t key_id, int32_t scheme, const vkzg_fr* d_f, const vkzg_g1_affine* C,
                                        const uint64_t* z, const vkzg_fr* y, const uint64_t* m_each, uint64_t K, vkzg_g1_affine* D,
                                        vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* yout);
/* ---- P2: verify_multiproof (multiproof.rs:178-215), IPA scheme; *ok = 1/0.  (KZG verification is two
 *      pairings, kzg/mod.rs:165-189, and stays on the host side of the shim.)                            */
int32_t vkzg_multiproof_verify_ipa(vkzg_ctx* ctx, uint32_t key_id, const vkzg_g1_affine* C, const uint64_t* z,
                                   const vkzg_fr* y, uint64_t m, const vkzg_g1_affine* D, const vkzg_g1_affine* L,
                                   const vkzg_g1_affine* R, const vkzg_fr* tip, const vkzg_fr* yproof, int32_t* ok);

/* ---- T1: Node::gen_commitment (verkle-tree/src/node.rs:212-277), level-synchronous -------------------- */
/* One level of node commitments in sparse (CSR) form.  Node j of the level commits to the terms
 * [row_ptr[j], row_ptr[j+1]) : term t places at slot slot[t] either the literal scalar lit[t] (child[t] < 0)
 * or to_data_item(nodes[child[t]]), where `nodes` is the array of ALL node commitments computed so far
 * (child ids are global; a node's children live in earlier levels).  d_out receives this level's n_nodes
 * commitments (normally d_nodes + the number of nodes of the earlier levels).                              */
int32_t vkzg_tree_level_dev(vkzg_ctx* ctx, uint32_t key_id, const uint32_t* d_row_ptr, uint64_t n_nodes,
                            const uint16_t* d_slot, const int32_t* d_child, const vkzg_fr* d_lit, uint64_t n_terms,
                            const vkzg_g1_affine* d_nodes, vkzg_g1_affine* d_out);
/* whole tree from host arrays: levels leaves-first, the last level is the single root node.               */
int32_t vkzg_tree_commit_levels(vkzg_ctx* ctx, uint32_t key_id, uint32_t n_levels, const uint64_t* nodes_per_level,
                                const uint32_t* const* row_ptr, const uint16_t* const* slot, const int32_t* const* child,
                                const vkzg_fr* const* lit, vkzg_g1_affine* root_out);

/* ---- next row: the verkle tree's host structure, VerkleTree::{new, insert_single, get_single, commitment}
 *      (verkle-tree/src/lib.rs:106-137) with Node::insert (node.rs:133-197) mirrored literally (Unit = u8).  Node
 *      commitments are cached and cleared along insertion paths like the reference; vkzg_tree_commit recommits only the
 *      dirty nodes, level by level.  ext_width = the reference's const generic N of the extension layout (quirk Q6), 256 for
 *      the Ethereum layout.  The key must have >= max(256, ext_width) bases.                                             */
typedef struct vkzg_tree vkzg_tree;
int32_t vkzg_tree_create(vkzg_tree** out, uint32_t key_len, uint32_t ext_width);
int32_t vkzg_tree_destroy(vkzg_tree* tree);
/* n (key, 32-byte value) pairs inserted IN ORDER; VKZG_ERR_RANGE at the first pair the reference panics on ("Traversed to
 * extension node with differing stem"), *n_done = pairs inserted before it                                               */
int32_t vkzg_tree_insert(vkzg_tree* tree, const uint8_t* keys, const uint8_t* values, uint64_t n, uint64_t* n_done);
/* 1 = found (value copied to value_out[32]), 0 = absent */
int32_t vkzg_tree_get(const vkzg_tree* tree, const uint8_t* key, uint8_t* value_out);
uint64_t vkzg_tree_nodes(const vkzg_tree* tree);
/* VerkleTree::path_to_stem (verkle-tree/src/lib.rs:131-137, node.rs:101-119): the internal nodes on the way to `stem`
 * (key_len bytes).  Entry d of the path is the internal node at depth d: its id, the unit stem[d] under which the walk
 * leaves it (the reference's prefix is stem[0..d+1]) and — when `commitments` is given — its cached commitment (`clean[d]`
 * = 1) or the identity (`clean[d]` = 0: the node is dirty, call vkzg_tree_commit first).  The walk stops at the first
 * extension node (not part of the path, like the reference).  VKZG_ERR_RANGE = VerkleError::InvalidPath (an internal node
 * has no child for the stem).  Arrays hold up to key_len entries; `commitments` / `clean` may be NULL.                  */
int32_t vkzg_tree_path_to_stem(const vkzg_tree* tree, const uint8_t* stem, uint32_t* path_len, uint32_t* node_ids, uint8_t* units,
                               vkzg_g1_affine* commitments, uint8_t* clean);
/* root commitment (Node::gen_commitment, node.rs:212-277) of the current tree; *n_committed = node commitments recomputed by
 * this call (0 when everything was cached).  Dirty extensions travel to the device as 65-byte records and their leaf-side
 * rows are expanded there; all new commitments are cached back on the host.  One tree must not be committed from two
 * threads at once (it is the caller's structure, like the reference's &mut self).                                     */
int32_t vkzg_tree_commit(vkzg_ctx* ctx, uint32_t key_id, vkzg_tree* tree, vkzg_g1_affine* root_out, uint64_t* n_committed);

/* ---- next row: KZG::setup (kzg/mod.rs:115-124) -------------------------------------------------------- */
/* powers[m] = [tau^i]G  ->  lagrange[n], n = next_pow2(m): the group inverse FFT over the radix-2 domain of size n
 * (`domain.ifft(&g1_points)`, inputs beyond m are the identity).                                                  */
int32_t vkzg_kzg_setup(vkzg_ctx* ctx, const vkzg_g1_affine* powers, uint32_t m, vkzg_g1_affine* lagrange);
int32_t vkzg_kzg_setup_dev(vkzg_ctx* ctx, const vkzg_g1_affine* d_powers, uint32_t m, vkzg_g1_affine* d_lagrange);
/* KZG::setup(max_items = m, gen) when the generator's secret tau is known to the caller (kzg/mod.rs:115-124 reads
 * gen.secret() itself): the same n = next_pow2(m) Lagrange points as vkzg_kzg_powers followed by vkzg_kzg_setup — canonical
 * affine, bit-identical — computed as n closed-form scalars times G (key_id = a window key whose base 0 is G).        */
int32_t vkzg_kzg_setup_from_secret(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* tau, uint32_t m, vkzg_g1_affine* lagrange);
/* KZGRandomPointGenerator::gen (kzg_point_generator.rs:32-43): out[i] = tau^i * G, i < m; key_id = a window key whose
 * base 0 is the generator G                                                                                       */
int32_t vkzg_kzg_powers(vkzg_ctx* ctx, uint32_t key_id, const vkzg_fr* tau, uint32_t m, vkzg_g1_affine* out);

/* ---- next row: IPAPointGenerator (vector-commit/src/ipa/ipa_point_generator.rs:51-81) with
 *      EthereumHashToCurve (:97-109): candidate i = SHA-256(seed || i as 8 little-endian bytes) parsed by ark-ec's
 *      Affine::from_random_bytes.  `gen(num)`: the first num indices 0, 1, 2 ... that give a point, in index order ->
 *      out[num]; *next_index (may be NULL) = the first index not consumed.  The caller enforces `num <= max`
 *      (PointGeneratorError::OutOfBounds) — the library has no notion of the generator's `max`.                          */
int32_t vkzg_ipa_crs_generate(vkzg_ctx* ctx, const uint8_t* seed, uint64_t seed_len, uint64_t num, vkzg_g1_affine* out,
                              uint64_t* next_index);
/* `gen_at(index)`: *ok = 1 and the point, or *ok = 0 (PointGeneratorError::InvalidPoint; out = identity encoding) */
int32_t vkzg_ipa_crs_generate_at(vkzg_ctx* ctx, const uint8_t* seed, uint64_t seed_len, uint64_t index, vkzg_g1_affine* out,
                                 int32_t* ok);

/* ---- several GPUs of one box behind the same boundary ------------------------------------------------
 * A group owns one vkzg_ctx per device and one host thread per device per call; ONE host process (e.g. the Rust caller of
 * the vector-commit traits) drives the box.  Width-N keys are replicated on every device and batches are cut into
 * contiguous ranges with no exchange (independent vectors); MSM keys are point-range sharded and the per-device partial
 * sums (64 bytes each) travel to device 0 over NVLink peer copies and are added there.  Results are byte-identical to the
 * single-device calls (canonical affine points).  device_ids == NULL: devices 0 .. ngpu-1 (ngpu == 0: every visible device);
 * an id may repeat (several contexts on one device — what the single-GPU tests use).
 * (One process per GPU instead — torchrun, NCCL — is verkle_kzg_b200/sharding.py: all_gather + vkzg_g1_sum_dev.)        */
typedef struct vkzg_mgpu vkzg_mgpu;
int32_t vkzg_mgpu_create(vkzg_mgpu** out, const int32_t* device_ids, uint32_t ngpu);
int32_t vkzg_mgpu_destroy(vkzg_mgpu* mg);
uint32_t vkzg_mgpu_size(const vkzg_mgpu* mg);
/* the i-th device's context, for the single-device entry points (keys loaded through it are its own) */
vkzg_ctx* vkzg_mgpu_ctx(vkzg_mgpu* mg, uint32_t i);
int32_t vkzg_mgpu_key_load(vkzg_mgpu* mg, const vkzg_g1_affine* bases, uint32_t n, const vkzg_g1_affine* q, uint32_t kind,
                           uint32_t window_bits, uint32_t* key_id);
int32_t vkzg_mgpu_key_free(vkzg_mgpu* mg, uint32_t key_id);
/* utils::inner_product as one MSM over all devices (KZG::commit at large n, configs[3]) */
int32_t vkzg_mgpu_msm(vkzg_mgpu* mg, uint32_t key_id, const vkzg_fr* scalars, uint64_t n, vkzg_g1_affine* out);
/* vkzg_commit_batch / vkzg_ipa_commit_prove_batch with the batch spread over the devices */
int32_t vkzg_mgpu_commit_batch(vkzg_mgpu* mg, uint32_t key_id, const vkzg_fr* scalars, uint32_t w, uint64_t B, vkzg_g1_affine* out);
int32_t vkzg_mgpu_ipa_commit_prove_batch(vkzg_mgpu* mg, uint32_t key_id, const vkzg_fr* a, const vkzg_fr* points, uint64_t B,
                                         vkzg_g1_affine* commitments, vkzg_g1_affine* L, vkzg_g1_affine* R, vkzg_fr* tip, vkzg_fr* y);

/* ---- measurement helpers (tools/, bench.py) ------------------------------------------------------------ */
/* enable != 0: bracket every launch of the dominant kernel (k_fixed_base_msm for window keys, k_msm_bucket for
 * MSM keys) with a CUDA event pair on the context's stream; calling it again clears the record.            */
int32_t vkzg_ctx_kernel_timing(vkzg_ctx* ctx, int32_t enable);
/* synchronises, then returns the number of bracketed launches and the sum of their durations */
int32_t vkzg_ctx_kernel_timing_read(vkzg_ctx* ctx, uint64_t* launches, double* total_ms);
/* runs `iters` dependent Fq multiplications on each of n elements in place: the integer-pipe probe */
int32_t vkzg_probe_fq_mul_dev(vkzg_ctx* ctx, vkzg_fq* d_x, const vkzg_fq* d_y, uint64_t n, uint32_t iters);
/* x <- x^(2^iters) (Montgomery form; any raw value <= 2p in, canonical out) on each of n elements in place through the hot
 * loops' multipliers: mode 0 = the dedicated square, 1 = the general product on equal operands, 2 = the Karatsuba product on
 * equal operands; modes 3 / 4: x <- x b^iters with b = x's 128-bit halves swapped and bits 254, 255 cleared, through the general /
 * the Karatsuba product (csrc/field.cuh, csrc/field_kara.cuh) */
int32_t vkzg_probe_fq_sqr_dev(vkzg_ctx* ctx, vkzg_fq* d_x, uint64_t n, uint32_t iters, uint32_t mode);
/* independent 32x32+64 multiply-accumulate chains on every SM; returns MAC32 issued (kind: 0 = mad.wide.u32,
 * 1 = mad.lo.u32, 2 = mad.wide + carry chain, 3 = fma.rn.f64) */
int32_t vkzg_probe_imad_dev(vkzg_ctx* ctx, uint32_t kind, uint32_t blocks, uint32_t threads, uint32_t iters,
                            uint64_t* macs_out);

#ifdef __cplusplus
}
#endif
#endif
