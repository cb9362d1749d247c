//! Raw bindings of libvkzg.so (include/vkzg.h).  Layout notes:
//!  * `ark_bn254::Fr` / `Fq` are `Fp<MontBackend<_, 4>>`: four little-endian u64 limbs in Montgomery form with
//!    R = 2^256 — the same 32 bytes as `vkzg_fr` (eight little-endian u32 limbs), so slices are passed as is.
//!  * points cross as affine `x || y` (64 bytes); the identity is all zeroes.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct vkzg_ctx {
    _private: [u8; 0],
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vkzg_fr {
    pub l: [u32; 8],
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vkzg_g1_affine {
    pub x: [u32; 8],
    pub y: [u32; 8],
}

pub const VKZG_KEY_WINDOW: u32 = 1;
pub const VKZG_KEY_MSM: u32 = 2;
pub const VKZG_ERR_RANGE: i32 = -3;

#[link(name = "vkzg")]
extern "C" {
    pub fn vkzg_strerror(status: i32) -> *const c_char;
    pub fn vkzg_ctx_create(out: *mut *mut vkzg_ctx, device_id: i32) -> i32;
    pub fn vkzg_ctx_destroy(ctx: *mut vkzg_ctx) -> i32;
    pub fn vkzg_key_load(ctx: *mut vkzg_ctx, bases: *const vkzg_g1_affine, n: u32, q: *const vkzg_g1_affine, kind: u32,
                         window_bits: u32, key_id: *mut u32) -> i32;
    pub fn vkzg_key_free(ctx: *mut vkzg_ctx, key_id: u32) -> i32;
    pub fn vkzg_msm(ctx: *mut vkzg_ctx, key_id: u32, scalars: *const vkzg_fr, n: u64, out: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_commit_batch(ctx: *mut vkzg_ctx, key_id: u32, scalars: *const vkzg_fr, w: u32, b: u64, out: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_to_data_item(ctx: *mut vkzg_ctx, points: *const vkzg_g1_affine, n: u64, out: *mut vkzg_fr) -> i32;
    pub fn vkzg_kzg_open_batch(ctx: *mut vkzg_ctx, key_id: u32, f: *const vkzg_fr, len: u32, domain_n: u32, points: *const vkzg_fr,
                               b: u64, proof: *mut vkzg_g1_affine, y: *mut vkzg_fr) -> i32;
    pub fn vkzg_kzg_prove_all_batch(ctx: *mut vkzg_ctx, key_id: u32, f: *const vkzg_fr, len: u32, domain_n: u32, b: u64,
                                    proof: *mut vkzg_g1_affine, y: *mut vkzg_fr) -> i32;
    pub fn vkzg_kzg_commit_open_batch(ctx: *mut vkzg_ctx, key_id: u32, f: *const vkzg_fr, len: u32, domain_n: u32, points: *const vkzg_fr,
                                      b: u64, commitments: *mut vkzg_g1_affine, proof: *mut vkzg_g1_affine, y: *mut vkzg_fr) -> i32;
    pub fn vkzg_ipa_prove_batch(ctx: *mut vkzg_ctx, key_id: u32, a: *const vkzg_fr, points: *const vkzg_fr,
                                commitments: *const vkzg_g1_affine, b: u64, prefix: *const u8, prefix_len: u32, dst: *const c_char,
                                l: *mut vkzg_g1_affine, r: *mut vkzg_g1_affine, tip: *mut vkzg_fr, y: *mut vkzg_fr) -> i32;
    pub fn vkzg_ipa_verify_batch(ctx: *mut vkzg_ctx, key_id: u32, points: *const vkzg_fr, commitments: *const vkzg_g1_affine, b: u64,
                                 prefix: *const u8, prefix_len: u32, dst: *const c_char, l: *const vkzg_g1_affine,
                                 r: *const vkzg_g1_affine, tip: *const vkzg_fr, y: *const vkzg_fr, ok: *mut i32) -> i32;
    pub fn vkzg_multiproof_prove(ctx: *mut vkzg_ctx, key_id: u32, scheme: i32, f: *const vkzg_fr, c: *const vkzg_g1_affine,
                                 z: *const u64, y: *const vkzg_fr, m: u64, d: *mut vkzg_g1_affine, l: *mut vkzg_g1_affine,
                                 r: *mut vkzg_g1_affine, tip: *mut vkzg_fr, yout: *mut vkzg_fr) -> i32;
    pub fn vkzg_multiproof_verify_ipa(ctx: *mut vkzg_ctx, key_id: u32, c: *const vkzg_g1_affine, z: *const u64, y: *const vkzg_fr, m: u64,
                                      d: *const vkzg_g1_affine, l: *const vkzg_g1_affine, r: *const vkzg_g1_affine,
                                      tip: *const vkzg_fr, yproof: *const vkzg_fr, ok: *mut i32) -> i32;
    pub fn vkzg_tree_commit_levels(ctx: *mut vkzg_ctx, key_id: u32, n_levels: u32, nodes_per_level: *const u64,
                                   row_ptr: *const *const u32, slot: *const *const u16, child: *const *const i32,
                                   lit: *const *const vkzg_fr, root_out: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_ipa_commit_prove_batch(ctx: *mut vkzg_ctx, key_id: u32, a: *const vkzg_fr, points: *const vkzg_fr, b: u64,
                                       commitments: *mut vkzg_g1_affine, l: *mut vkzg_g1_affine, r: *mut vkzg_g1_affine,
                                       tip: *mut vkzg_fr, y: *mut vkzg_fr) -> i32;
    pub fn vkzg_kzg_setup(ctx: *mut vkzg_ctx, powers: *const vkzg_g1_affine, m: u32, lagrange: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_tree_create(out: *mut *mut Opaque, key_len: u32, ext_width: u32) -> i32;
    pub fn vkzg_tree_insert(tree: *mut Opaque, keys: *const u8, values: *const u8, n: u64, n_done: *mut u64) -> i32;
    pub fn vkzg_tree_get(tree: *const Opaque, key: *const u8, value_out: *mut u8) -> i32;
    pub fn vkzg_tree_commit(ctx: *mut vkzg_ctx, key_id: u32, tree: *mut Opaque, root_out: *mut vkzg_g1_affine, n_committed: *mut u64) -> i32;
    pub fn vkzg_tree_destroy(tree: *mut Opaque) -> i32;
    pub fn vkzg_ctx_sync(ctx: *mut vkzg_ctx) -> i32;
    pub fn vkzg_ctx_trim(ctx: *mut vkzg_ctx) -> i32;
    pub fn vkzg_ipa_prove_commitment_batch(ctx: *mut vkzg_ctx, key_id: u32, a: *const vkzg_fr, commitments: *const vkzg_g1_affine, b: u64,
                                           l: *mut vkzg_g1_affine, r: *mut vkzg_g1_affine, tip: *mut vkzg_fr) -> i32;
    pub fn vkzg_ipa_verify_commitment_batch(ctx: *mut vkzg_ctx, key_id: u32, commitments: *const vkzg_g1_affine, b: u64,
                                            l: *const vkzg_g1_affine, r: *const vkzg_g1_affine, tip: *const vkzg_fr, ok: *mut i32) -> i32;
    pub fn vkzg_tree_path_to_stem(tree: *const Opaque, stem: *const u8, path_len: *mut u32, node_ids: *mut u32, units: *mut u8,
                                  commitments: *mut vkzg_g1_affine, clean: *mut u8) -> i32;
    pub fn vkzg_multiproof_prove_batch(ctx: *mut vkzg_ctx, key_id: u32, scheme: i32, f: *const vkzg_fr, c: *const vkzg_g1_affine,
                                       z: *const u64, y: *const vkzg_fr, m_each: *const u64, k: u64, d: *mut vkzg_g1_affine,
                                       l: *mut vkzg_g1_affine, r: *mut vkzg_g1_affine, tip: *mut vkzg_fr, yout: *mut vkzg_fr) -> i32;
    // setup code on the device
    pub fn vkzg_kzg_powers(ctx: *mut vkzg_ctx, key_id: u32, tau: *const vkzg_fr, m: u32, out: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_kzg_setup_from_secret(ctx: *mut vkzg_ctx, key_id: u32, tau: *const vkzg_fr, m: u32, lagrange: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_ipa_crs_generate(ctx: *mut vkzg_ctx, seed: *const u8, seed_len: u64, num: u64, out: *mut vkzg_g1_affine,
                                 next_index: *mut u64) -> i32;
    pub fn vkzg_ipa_crs_generate_at(ctx: *mut vkzg_ctx, seed: *const u8, seed_len: u64, index: u64, out: *mut vkzg_g1_affine,
                                    ok: *mut i32) -> i32;
    // several GPUs of one box behind the same boundary (one host process)
    pub fn vkzg_mgpu_create(out: *mut *mut Opaque, device_ids: *const i32, ngpu: u32) -> i32;
    pub fn vkzg_mgpu_destroy(mg: *mut Opaque) -> i32;
    pub fn vkzg_mgpu_size(mg: *const Opaque) -> u32;
    pub fn vkzg_mgpu_ctx(mg: *mut Opaque, i: u32) -> *mut vkzg_ctx;
    pub fn vkzg_mgpu_key_load(mg: *mut Opaque, bases: *const vkzg_g1_affine, n: u32, q: *const vkzg_g1_affine, kind: u32,
                              window_bits: u32, key_id: *mut u32) -> i32;
    pub fn vkzg_mgpu_key_free(mg: *mut Opaque, key_id: u32) -> i32;
    pub fn vkzg_mgpu_msm(mg: *mut Opaque, key_id: u32, scalars: *const vkzg_fr, n: u64, out: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_mgpu_commit_batch(mg: *mut Opaque, key_id: u32, scalars: *const vkzg_fr, w: u32, b: u64, out: *mut vkzg_g1_affine) -> i32;
    pub fn vkzg_mgpu_ipa_commit_prove_batch(mg: *mut Opaque, key_id: u32, a: *const vkzg_fr, points: *const vkzg_fr, b: u64,
                                            commitments: *mut vkzg_g1_affine, l: *mut vkzg_g1_affine, r: *mut vkzg_g1_affine,
                                            tip: *mut vkzg_fr, y: *mut vkzg_fr) -> i32;
}

pub type Opaque = c_void;
