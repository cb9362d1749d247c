//! `GpuIpa<N>` / `GpuKzg`: the reference's `VectorCommitment` + `VectorCommitmentMultiproof` traits over libvkzg.
//!
//! This file is meant to live INSIDE the `vector-commit` crate as `src/gpu.rs` behind `--features b200`
//! (proof fields, `TranscriptHasher` internals and `utils` are crate-private in the reference:
//! lib.rs:24-25, ipa/mod.rs:79-84, kzg/mod.rs:81-84, multiproof.rs:55-58).  It cannot be compiled in the
//! build container (no cargo/rustc); the identical C ABI is exercised from Python by the test-suite.
use std::ffi::CString;
use std::marker::PhantomData;
use std::ptr;

use ark_bn254::{Bn254, Fr, G1Affine, G1Projective};
use ark_ec::{AffineRepr, CurveGroup};
use ark_ff::{field_hashers::DefaultFieldHasher, Zero};
use ark_poly::GeneralEvaluationDomain;
use sha2::Sha256;

use crate::ffi::*;
use crate::ipa::{IPACommitProof, IPAError, IPAProof, IPAUniversalParams};
use crate::kzg::{KZGError, KZGKey, KZGProof};
use crate::lagrange_basis::LagrangeBasis;
use crate::multiproof::{Multiproof, MultiproofProverQuery, MultiproofVerifierQuery, VectorCommitmentMultiproof};
use crate::transcript::TranscriptHasher;
use crate::{VCUniversalParams, VectorCommitment};

type D = GeneralEvaluationDomain<Fr>;
type H = DefaultFieldHasher<Sha256>;

/// One GPU context + one resident key (window tables) per `UniversalParams`.
pub struct GpuKey {
    ctx: *mut vkzg_ctx,
    id: u32,
    pub size: usize,
}
unsafe impl Sync for GpuKey {} // calls are serialised by &mut-free associated fns on one thread

fn to_abi(p: &G1Projective) -> vkzg_g1_affine {
    let a: G1Affine = p.into_affine();
    if a.is_zero() {
        return vkzg_g1_affine::default();
    }
    // Fq is four Montgomery u64 limbs: reinterpret as eight u32 (little-endian hosts)
    unsafe { vkzg_g1_affine { x: std::mem::transmute(a.x.0 .0), y: std::mem::transmute(a.y.0 .0) } }
}
fn from_abi(p: &vkzg_g1_affine) -> G1Projective {
    if p.x.iter().chain(p.y.iter()).all(|w| *w == 0) {
        return G1Projective::zero();
    }
    unsafe {
        let x = ark_bn254::Fq::new_unchecked(ark_ff::BigInt(std::mem::transmute(p.x)));
        let y = ark_bn254::Fq::new_unchecked(ark_ff::BigInt(std::mem::transmute(p.y)));
        G1Affine::new_unchecked(x, y).into()
    }
}
fn fr_ptr(v: &[Fr]) -> *const vkzg_fr {
    v.as_ptr() as *const vkzg_fr
}

impl GpuKey {
    pub fn load(g: &[G1Projective], q: Option<&G1Projective>) -> Self {
        let mut ctx = ptr::null_mut();
        assert_eq!(unsafe { vkzg_ctx_create(&mut ctx, 0) }, 0, "no sm_100 GPU: libvkzg has no CPU fallback");
        let bases: Vec<vkzg_g1_affine> = g.iter().map(to_abi).collect();
        let qa = q.map(to_abi);
        let mut id = 0u32;
        let st = unsafe {
            vkzg_key_load(ctx, bases.as_ptr(), bases.len() as u32, qa.as_ref().map_or(ptr::null(), |p| p as *const _), VKZG_KEY_WINDOW, 0, &mut id)
        };
        assert_eq!(st, 0);
        GpuKey { ctx, id, size: g.len() }
    }
}
impl Drop for GpuKey {
    fn drop(&mut self) {
        unsafe {
            vkzg_key_free(self.ctx, self.id);
            vkzg_ctx_destroy(self.ctx);
        }
    }
}

/// IPA over the GPU.  `UniversalParams` wraps the reference's params (kept for `verify` fallbacks and for
/// callers that read `g`/`q`) plus the resident GPU key.
pub struct GpuIpaParams<const N: usize> {
    pub inner: IPAUniversalParams<N, G1Projective, H>,
    pub gpu: GpuKey,
}
impl<const N: usize> VCUniversalParams for GpuIpaParams<N> {
    fn max_size(&self) -> usize {
        N
    }
}

pub struct GpuIpa<const N: usize> {
    _p: PhantomData<[(); N]>,
}

impl<const N: usize> VectorCommitment for GpuIpa<N> {
    type UniversalParams = GpuIpaParams<N>;
    type Commitment = G1Projective;
    type Data = LagrangeBasis<Fr, D>;
    type Proof = IPAProof<G1Projective>;
    type BatchProof = Vec<Self::Proof>;
    type Error = IPAError;
    type PointGenerator = crate::ipa::ipa_point_generator::IPAPointGenerator<G1Projective, crate::ipa::ipa_point_generator::EthereumHashToCurve>;
    type Transcript = TranscriptHasher<Fr, H>;

    fn setup(max_items: usize, gen: &Self::PointGenerator) -> Result<Self::UniversalParams, crate::PointGeneratorError> {
        let inner = crate::ipa::IPA::<N, G1Projective, H, D>::setup(max_items, gen)?; // ipa/mod.rs:121-128 (setup is not the hot path)
        let gpu = GpuKey::load(&inner.g, Some(&inner.q));
        Ok(GpuIpaParams { inner, gpu })
    }

    /// ipa/mod.rs:130-135 -> vkzg_commit_batch (B = 1).  Batch callers should use `commit_many`.
    fn commit(key: &Self::UniversalParams, data: &Self::Data) -> Result<Self::Commitment, Self::Error> {
        let ev = data.elements_ref();
        let w = ev.len().min(key.gpu.size); // inner_product zips (utils.rs:17)
        let mut out = vkzg_g1_affine::default();
        let st = unsafe { vkzg_commit_batch(key.gpu.ctx, key.gpu.id, fr_ptr(ev), w as u32, 1, &mut out) };
        if st != 0 {
            return Err(IPAError::OutOfCRS);
        }
        Ok(from_abi(&out))
    }

    /// ipa/mod.rs:137-154 + low_level_ipa :268-319 -> vkzg_ipa_prove_batch (B = 1)
    fn prove_point(key: &Self::UniversalParams, commitment: &Self::Commitment, point: Fr, data: &Self::Data,
                   transcript: Option<Self::Transcript>) -> Result<Self::Proof, Self::Error> {
        let rounds = N.trailing_zeros() as usize;
        let (prefix, dst) = match &transcript {
            Some(t) => (t.state_bytes().to_vec(), t.domain_label().to_vec()), // crate-private accessors added by the patch
            None => (Vec::new(), b"ipa".to_vec()),
        };
        let dst = CString::new(dst).unwrap();
        let c = to_abi(commitment);
        let (mut l, mut r) = (vec![vkzg_g1_affine::default(); rounds], vec![vkzg_g1_affine::default(); rounds]);
        let (mut tip, mut y) = (Fr::zero(), Fr::zero());
        // the library reads a[B][N]: a shorter evaluation vector (LagrangeBasis::from_vec of < N items over a smaller
        // domain) must not be handed over as is.  The reference zips a with b and G (utils.rs:17, ipa/mod.rs:281-283), i.e.
        // it would silently prove a truncated statement; here the width must match the key (IPAError::OutOfCRS otherwise).
        if data.elements_ref().len() != N {
            return Err(IPAError::OutOfCRS);
        }
        let st = unsafe {
            vkzg_ipa_prove_batch(key.gpu.ctx, key.gpu.id, fr_ptr(data.elements_ref()), &point as *const Fr as *const vkzg_fr, &c, 1,
                                 prefix.as_ptr(), prefix.len() as u32, dst.as_ptr(), l.as_mut_ptr(), r.as_mut_ptr(),
                                 &mut tip as *mut Fr as *mut vkzg_fr, &mut y as *mut Fr as *mut vkzg_fr)
        };
        if st != 0 {
            return Err(IPAError::OutOfDomain);
        }
        Ok(IPAProof { l: l.iter().map(from_abi).collect(), r: r.iter().map(from_abi).collect(), tip, y })
    }

    fn prove_batch(_: &Self::UniversalParams, _: &Self::Commitment, _: Vec<usize>, _: &Self::Data) -> Result<Self::BatchProof, Self::Error> {
        todo!() // as the reference (ipa/mod.rs:156-163)
    }

    /// ipa/mod.rs:165-181 + :321-360 -> vkzg_ipa_verify_batch (B = 1)
    fn verify_point(key: &Self::UniversalParams, commitment: &Self::Commitment, point: Fr, proof: &Self::Proof,
                    transcript: Option<Self::Transcript>) -> Result<bool, Self::Error> {
        let (prefix, dst) = match &transcript {
            Some(t) => (t.state_bytes().to_vec(), t.domain_label().to_vec()),
            None => (Vec::new(), b"ipa".to_vec()),
        };
        let dst = CString::new(dst).unwrap();
        let c = to_abi(commitment);
        let l: Vec<_> = proof.l.iter().map(to_abi).collect();
        let r: Vec<_> = proof.r.iter().map(to_abi).collect();
        let mut ok = 0i32;
        let st = unsafe {
            vkzg_ipa_verify_batch(key.gpu.ctx, key.gpu.id, &point as *const Fr as *const vkzg_fr, &c, 1, prefix.as_ptr(), prefix.len() as u32,
                                  dst.as_ptr(), l.as_ptr(), r.as_ptr(), &proof.tip as *const Fr as *const vkzg_fr,
                                  &proof.y as *const Fr as *const vkzg_fr, &mut ok)
        };
        if st != 0 {
            return Err(IPAError::OutOfDomain);
        }
        Ok(ok == 1)
    }

    fn verify_batch(_: &Self::UniversalParams, _: &Self::Commitment, _: &Self::BatchProof) -> Result<bool, Self::Error> {
        todo!()
    }
}

impl<const N: usize> GpuIpa<N> {
    /// ipa/mod.rs:199-235 -> vkzg_ipa_prove_commitment_batch (B = 1)
    pub fn prove_commitment(key: &GpuIpaParams<N>, commitment: &G1Projective, data: &LagrangeBasis<Fr, D>) -> Result<IPACommitProof<G1Projective>, IPAError> {
        if data.elements_ref().len() != N {
            return Err(IPAError::OutOfCRS); // (the library reads a[B][N])
        }
        let rounds = N.trailing_zeros() as usize;
        let c = to_abi(commitment);
        let (mut l, mut r) = (vec![vkzg_g1_affine::default(); rounds], vec![vkzg_g1_affine::default(); rounds]);
        let mut tip = Fr::zero();
        let st = unsafe {
            vkzg_ipa_prove_commitment_batch(key.gpu.ctx, key.gpu.id, fr_ptr(data.elements_ref()), &c, 1, l.as_mut_ptr(), r.as_mut_ptr(),
                                            &mut tip as *mut Fr as *mut vkzg_fr)
        };
        if st != 0 {
            return Err(IPAError::OutOfDomain);
        }
        Ok(IPACommitProof { l: l.iter().map(from_abi).collect(), r: r.iter().map(from_abi).collect(), tip })
    }

    /// ipa/mod.rs:238-265 -> vkzg_ipa_verify_commitment_batch (B = 1)
    pub fn verify_commitment_proof(key: &GpuIpaParams<N>, commitment: &G1Projective, proof: &IPACommitProof<G1Projective>) -> bool {
        if proof.l.len() != N.trailing_zeros() as usize || proof.r.len() != proof.l.len() {
            return false; // (the reference sizes gens by the proof; the library by the key)
        }
        let c = to_abi(commitment);
        let l: Vec<_> = proof.l.iter().map(to_abi).collect();
        let r: Vec<_> = proof.r.iter().map(to_abi).collect();
        let mut ok = 0i32;
        let st = unsafe {
            vkzg_ipa_verify_commitment_batch(key.gpu.ctx, key.gpu.id, &c, 1, l.as_ptr(), r.as_ptr(), &proof.tip as *const Fr as *const vkzg_fr, &mut ok)
        };
        st == 0 && ok == 1
    }

    /// B commitments in one launch — what `Node::gen_commitment` and bulk callers should use.
    pub fn commit_many(key: &GpuIpaParams<N>, rows: &[Fr], width: usize) -> Vec<G1Projective> {
        assert!(width > 0 && rows.len() % width == 0, "rows must hold whole vectors");
        let b = rows.len() / width;
        let mut out = vec![vkzg_g1_affine::default(); b];
        let st = unsafe { vkzg_commit_batch(key.gpu.ctx, key.gpu.id, fr_ptr(rows), width as u32, b as u64, out.as_mut_ptr()) };
        assert_eq!(st, 0);
        out.iter().map(from_abi).collect()
    }
}

/// multiproof.rs:218-225 has an empty impl (the default methods do the work); here the default
/// `prove_multiproof` / `verify_multiproof` are overridden with one library call each.
impl<const N: usize> VectorCommitmentMultiproof for GpuIpa<N> {
    fn prove_multiproof<'a>(key: &Self::UniversalParams, queries: impl Iterator<Item = &'a MultiproofProverQuery<'a, Fr, G1Projective, Self::Data>>)
        -> Result<Multiproof<Self::Proof, G1Projective>, Self::Error> {
        let qs: Vec<_> = queries.collect();
        let m = qs.len();
        let rounds = N.trailing_zeros() as usize;
        let mut f: Vec<Fr> = Vec::with_capacity(m * N);
        let (mut c, mut z, mut y) = (Vec::with_capacity(m), Vec::with_capacity(m), Vec::with_capacity(m));
        for q in &qs {
            if q.data.elements_ref().len() != N {
                return Err(IPAError::OutOfCRS); // rows are read as f[m][N]
            }
            f.extend_from_slice(q.data.elements_ref());
            c.push(to_abi(q.commit));
            z.push(q.z as u64);
            y.push(q.y);
        }
        let mut d = vkzg_g1_affine::default();
        let (mut l, mut r) = (vec![vkzg_g1_affine::default(); rounds], vec![vkzg_g1_affine::default(); rounds]);
        let (mut tip, mut yo) = (Fr::zero(), Fr::zero());
        let st = unsafe {
            vkzg_multiproof_prove(key.gpu.ctx, key.gpu.id, 0, fr_ptr(&f), c.as_ptr(), z.as_ptr(), fr_ptr(&y), m as u64, &mut d, l.as_mut_ptr(),
                                  r.as_mut_ptr(), &mut tip as *mut Fr as *mut vkzg_fr, &mut yo as *mut Fr as *mut vkzg_fr)
        };
        if st != 0 {
            return Err(IPAError::OutOfDomain);
        }
        Ok(Multiproof { proof: IPAProof { l: l.iter().map(from_abi).collect(), r: r.iter().map(from_abi).collect(), tip, y: yo }, d: from_abi(&d) })
    }

    fn verify_multiproof<'a>(key: &Self::UniversalParams, queries: impl Iterator<Item = &'a MultiproofVerifierQuery<'a, Fr, G1Projective>>,
                             proof: &Multiproof<Self::Proof, G1Projective>) -> Result<bool, Self::Error> {
        let (mut c, mut z, mut y) = (Vec::new(), Vec::new(), Vec::new());
        for q in queries {
            c.push(to_abi(q.commit));
            z.push(q.z as u64);
            y.push(q.y);
        }
        let l: Vec<_> = proof.proof.l.iter().map(to_abi).collect();
        let r: Vec<_> = proof.proof.r.iter().map(to_abi).collect();
        let d = to_abi(&proof.d);
        let mut ok = 0i32;
        let st = unsafe {
            vkzg_multiproof_verify_ipa(key.gpu.ctx, key.gpu.id, c.as_ptr(), z.as_ptr(), fr_ptr(&y), c.len() as u64, &d, l.as_ptr(), r.as_ptr(),
                                       &proof.proof.tip as *const Fr as *const vkzg_fr, &proof.proof.y as *const Fr as *const vkzg_fr, &mut ok)
        };
        if st != 0 {
            return Err(IPAError::OutOfDomain);
        }
        Ok(ok == 1)
    }
}

/// KZG over the GPU: commit and prove_point on the device, verify_point = the reference's two pairings on the host.
pub struct GpuKzgParams {
    pub inner: KZGKey<Fr, G1Projective, <Bn254 as ark_ec::pairing::Pairing>::G2>,
    pub gpu: GpuKey,
}
impl VCUniversalParams for GpuKzgParams {
    fn max_size(&self) -> usize {
        self.gpu.size
    }
}
pub struct GpuKzg;

impl GpuKzg {
    /// kzg/mod.rs:126-134
    pub fn commit(key: &GpuKzgParams, data: &LagrangeBasis<Fr, D>) -> Result<G1Projective, KZGError> {
        let ev = data.elements_ref();
        let w = ev.len().min(key.gpu.size);
        let mut out = vkzg_g1_affine::default();
        match unsafe { vkzg_commit_batch(key.gpu.ctx, key.gpu.id, fr_ptr(ev), w as u32, 1, &mut out) } {
            0 => Ok(from_abi(&out)),
            _ => Err(KZGError::OutOfCRS),
        }
    }
    /// kzg/mod.rs:136-154 -> vkzg_kzg_open_batch (B = 1); `domain_n` = the data's domain size (from_vec_and_domain)
    pub fn prove_point(key: &GpuKzgParams, point: Fr, data: &LagrangeBasis<Fr, D>) -> Result<KZGProof<Fr, G1Projective>, KZGError> {
        let ev = data.elements_ref();
        let mut pf = vkzg_g1_affine::default();
        let mut y = Fr::zero();
        let st = unsafe {
            vkzg_kzg_open_batch(key.gpu.ctx, key.gpu.id, fr_ptr(ev), ev.len() as u32, data.domain_size() as u32, &point as *const Fr as *const vkzg_fr,
                                1, &mut pf, &mut y as *mut Fr as *mut vkzg_fr)
        };
        match st {
            0 => Ok(KZGProof { proof: from_abi(&pf), y }),
            VKZG_ERR_RANGE => Err(KZGError::OutOfDomain), // the reference panics here (quirk Q2)
            _ => Err(KZGError::OutOfCRS),
        }
    }
    /// kzg/mod.rs:200-235 (`prove_all_points`, unreachable in the reference): one opening per point of the data's domain,
    /// entry i equal to `prove_point(key, i, data)` -> vkzg_kzg_prove_all_batch (B = 1)
    pub fn prove_all_points(key: &GpuKzgParams, data: &LagrangeBasis<Fr, D>) -> Result<Vec<KZGProof<Fr, G1Projective>>, KZGError> {
        let ev = data.elements_ref();
        let dn = data.domain_size().max(ev.len()).next_power_of_two();
        let mut pf = vec![vkzg_g1_affine::default(); dn];
        let mut y = vec![Fr::zero(); dn];
        let st = unsafe {
            vkzg_kzg_prove_all_batch(key.gpu.ctx, key.gpu.id, fr_ptr(ev), ev.len() as u32, data.domain_size() as u32, 1, pf.as_mut_ptr(),
                                     y.as_mut_ptr() as *mut vkzg_fr)
        };
        match st {
            0 => Ok(pf.iter().zip(y).map(|(p, y)| KZGProof { proof: from_abi(p), y }).collect()),
            _ => Err(KZGError::OutOfCRS),
        }
    }
    /// kzg/mod.rs:165-189: unchanged host pairing check
    pub fn verify_point(key: &GpuKzgParams, commitment: &G1Projective, point: Fr, proof: &KZGProof<Fr, G1Projective>) -> Result<bool, KZGError> {
        crate::kzg::KZG::<Bn254, H, D>::verify_point(&key.inner, commitment, point, proof, None)
    }
}

/// `IPAPointGenerator<G1Projective, EthereumHashToCurve>` (ipa/ipa_point_generator.rs:14-86) with the SHA-256
/// try-and-increment and the square roots on the device (`vkzg_ipa_crs_generate`): same constructor, bounds and errors.
pub struct GpuIpaPointGenerator {
    max: usize,
    seed: Vec<u8>,
}
impl GpuIpaPointGenerator {
    pub fn new(max: usize, seed: Vec<u8>) -> Self {
        Self { max, seed }
    }
    pub fn set_max(&mut self, max: usize) {
        self.max = max;
    }
    fn with_ctx<T>(f: impl FnOnce(*mut vkzg_ctx) -> T) -> T {
        let mut ctx = ptr::null_mut();
        assert_eq!(unsafe { vkzg_ctx_create(&mut ctx, 0) }, 0, "no sm_100 GPU: libvkzg has no CPU fallback");
        let r = f(ctx);
        unsafe { vkzg_ctx_destroy(ctx) };
        r
    }
}
impl Default for GpuIpaPointGenerator {
    fn default() -> Self {
        Self { max: 256, seed: "eth_verkle_oct_2021".to_owned().into_bytes() } // ipa_point_generator.rs:38-47
    }
}
impl crate::PointGenerator for GpuIpaPointGenerator {
    type Point = G1Projective;
    type Secret = Vec<u8>;

    /// ipa_point_generator.rs:51-70
    fn gen(&self, num: usize) -> Result<Vec<G1Projective>, crate::PointGeneratorError> {
        if num > self.max {
            return Err(crate::PointGeneratorError::OutOfBounds);
        }
        if num == 0 {
            return Ok(vec![]);
        }
        let mut out = vec![vkzg_g1_affine::default(); num];
        let st = Self::with_ctx(|ctx| unsafe {
            vkzg_ipa_crs_generate(ctx, self.seed.as_ptr(), self.seed.len() as u64, num as u64, out.as_mut_ptr(), ptr::null_mut())
        });
        assert_eq!(st, 0);
        Ok(out.iter().map(from_abi).collect())
    }
    /// ipa_point_generator.rs:72-81
    fn gen_at(&self, index: usize) -> Result<G1Projective, crate::PointGeneratorError> {
        if index > self.max {
            return Err(crate::PointGeneratorError::OutOfBounds);
        }
        let mut out = vkzg_g1_affine::default();
        let mut ok = 0i32;
        let st = Self::with_ctx(|ctx| unsafe {
            vkzg_ipa_crs_generate_at(ctx, self.seed.as_ptr(), self.seed.len() as u64, index as u64, &mut out, &mut ok)
        });
        assert_eq!(st, 0);
        if ok == 1 { Ok(from_abi(&out)) } else { Err(crate::PointGeneratorError::InvalidPoint) }
    }
    fn secret(&self) -> Option<Vec<u8>> {
        Some(self.seed.clone())
    }
}

/// All GPUs of the box from ONE host process: `vkzg_mgpu_*` owns a context and a host thread per device.
/// Width-N keys are replicated and a batch is cut into contiguous ranges (no exchange); an MSM key is point-range sharded
/// and the 64-byte partial sums are added on device 0 after NVLink peer copies.
pub struct GpuBox {
    mg: *mut Opaque,
}
impl GpuBox {
    /// `devices = &[]`: every visible GPU
    pub fn new(devices: &[i32]) -> Self {
        let mut mg = ptr::null_mut();
        let st = unsafe { vkzg_mgpu_create(&mut mg, if devices.is_empty() { ptr::null() } else { devices.as_ptr() }, devices.len() as u32) };
        assert_eq!(st, 0, "no sm_100 GPU: libvkzg has no CPU fallback");
        GpuBox { mg }
    }
    pub fn gpus(&self) -> u32 {
        unsafe { vkzg_mgpu_size(self.mg) }
    }
    pub fn load_window_key(&self, g: &[G1Projective], q: Option<&G1Projective>) -> u32 {
        let bases: Vec<vkzg_g1_affine> = g.iter().map(to_abi).collect();
        let qa = q.map(to_abi);
        let mut id = 0u32;
        let st = unsafe {
            vkzg_mgpu_key_load(self.mg, bases.as_ptr(), bases.len() as u32, qa.as_ref().map_or(ptr::null(), |p| p as *const _), VKZG_KEY_WINDOW, 0, &mut id)
        };
        assert_eq!(st, 0);
        id
    }
    pub fn load_msm_key(&self, g: &[G1Projective]) -> u32 {
        let bases: Vec<vkzg_g1_affine> = g.iter().map(to_abi).collect();
        let mut id = 0u32;
        assert_eq!(unsafe { vkzg_mgpu_key_load(self.mg, bases.as_ptr(), bases.len() as u32, ptr::null(), VKZG_KEY_MSM, 0, &mut id) }, 0);
        id
    }
    /// utils::inner_product (utils.rs:16-19) over all devices: KZG::commit at large n (configs[3])
    pub fn msm(&self, key: u32, scalars: &[Fr]) -> G1Projective {
        let mut out = vkzg_g1_affine::default();
        assert_eq!(unsafe { vkzg_mgpu_msm(self.mg, key, fr_ptr(scalars), scalars.len() as u64, &mut out) }, 0);
        from_abi(&out)
    }
    /// `rows.len() / width` independent commits spread over the devices
    pub fn commit_many(&self, key: u32, rows: &[Fr], width: usize) -> Vec<G1Projective> {
        let b = rows.len() / width;
        let mut out = vec![vkzg_g1_affine::default(); b];
        assert_eq!(unsafe { vkzg_mgpu_commit_batch(self.mg, key, fr_ptr(rows), width as u32, b as u64, out.as_mut_ptr()) }, 0);
        out.iter().map(from_abi).collect()
    }
}
impl Drop for GpuBox {
    fn drop(&mut self) {
        unsafe { vkzg_mgpu_destroy(self.mg) };
    }
}
