//! Golden-vector generator for verkle_kzg_b200 (https://… the B200 drop-in for this crate's commitment path).
//!
//! NOT part of the reference.  `make_vectors.py` (next to this file) copies it into a scratch copy of
//! `vector-commit/src/`, declares it in lib.rs as `#[cfg(test)] mod arkworks_vectors;`, widens the visibility of a
//! few proof-struct fields to `pub(crate)` in that scratch copy, runs
//!     cargo test -p vector-commit arkworks_vectors -- --nocapture
//! and stores the JSON printed between ARKVEC_BEGIN / ARKVEC_END as tests/golden/arkworks/vectors.json.
//! Every input is a closed formula (no RNG) that tests/test_arkworks_vectors.py rebuilds on its side; every output
//! is what arkworks 0.4 + this crate produce.  Encodings: points = ark-serialize compressed (32 bytes, hex) and
//! uncompressed x || y little-endian canonical (64 bytes, hex); scalars = 32 little-endian canonical bytes (hex).
use ark_bn254::Bn254;
use ark_ec::{pairing::Pairing, CurveGroup, Group};
use ark_ff::{
    field_hashers::{DefaultFieldHasher, HashToField},
    One, PrimeField, Zero,
};
use ark_poly::{EvaluationDomain, GeneralEvaluationDomain};
use ark_serialize::CanonicalSerialize;
use sha2::Sha256;

use crate::{
    ipa::{IPAPointGenerator, IPAUniversalParams, IPA},
    kzg::{kzg_point_generator::KZGRandomPointGenerator, KZG},
    lagrange_basis::LagrangeBasis,
    multiproof::{MultiproofProverQuery, VectorCommitmentMultiproof},
    transcript::{Transcript, TranscriptHasher},
    PointGenerator, VCCommitment, VectorCommitment,
};

type F = <Bn254 as Pairing>::ScalarField;
type G = <Bn254 as Pairing>::G1;
type Hasher = DefaultFieldHasher<Sha256>;
type Dom = GeneralEvaluationDomain<F>;
type T = TranscriptHasher<F, Hasher>;

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
fn fr(x: &F) -> String {
    let mut b = Vec::new();
    x.serialize_compressed(&mut b).unwrap();
    format!("\"{}\"", hex(&b))
}
fn pc(p: &G) -> String {
    let mut b = Vec::new();
    p.serialize_compressed(&mut b).unwrap();
    format!("\"{}\"", hex(&b))
}
/// uncompressed affine x || y (the identity serialises as ark-serialize does it: x = 0, y = 0 with the infinity flag)
fn pu(p: &G) -> String {
    let mut b = Vec::new();
    p.into_affine().serialize_uncompressed(&mut b).unwrap();
    format!("\"{}\"", hex(&b))
}
fn list(v: Vec<String>) -> String {
    format!("[{}]", v.join(","))
}
fn gens(n: usize) -> Vec<G> {
    // the bases of ipa/mod.rs:385-387: g_i = (i + 1) G, q = (n + 1) G
    (0..n + 1).map(|i| G::generator() * F::from(i as u64 + 1)).collect()
}
fn data(n: usize, a: u64, b: u64) -> Vec<F> {
    (0..n as u64).map(|i| F::from(a) + F::from(b) * F::from(i)).collect()
}

fn ipa_case<const N: usize>(out: &mut Vec<String>) {
    type I<const M: usize> = IPA<M, G, Hasher, Dom>;
    let key = IPAUniversalParams::<N, G, Hasher>::new_from_vec(gens(N));
    let d = LagrangeBasis::<F, Dom>::from_vec(data(N, 7, 3));
    let c = I::<N>::commit(&key, &d).unwrap();
    let mut proofs = Vec::new();
    for idx in [1usize, N - 1, N + 1, 2 * N] {
        let p = I::<N>::prove(&key, &c, idx, &d).unwrap();
        assert!(I::<N>::verify(&key, &c, idx, &p).unwrap());
        proofs.push(format!(
            "{{\"index\":{},\"l\":{},\"r\":{},\"tip\":{},\"y\":{}}}",
            idx,
            list(p.l.iter().map(pc).collect()),
            list(p.r.iter().map(pc).collect()),
            fr(&p.tip),
            fr(&p.y)
        ));
    }
    let cp = I::<N>::prove_commitment(&key, &c, &d);
    assert!(I::<N>::verify_commitment_proof(&key, &c, &cp));
    out.push(format!(
        "{{\"n\":{},\"data\":\"7+3i\",\"commit\":{},\"commit_xy\":{},\"to_data_item\":{},\"proofs\":{},\"commit_proof\":{{\"l\":{},\"r\":{},\"tip\":{}}}}}",
        N,
        pc(&c),
        pu(&c),
        fr(&c.to_data_item()),
        list(proofs),
        list(cp.l.iter().map(pc).collect()),
        list(cp.r.iter().map(pc).collect()),
        fr(&cp.tip)
    ));
}

#[test]
fn arkworks_vectors() {
    let g = G::generator();
    let mut j: Vec<String> = Vec::new();

    // 1. compressed-point flags (transcript.rs:64-71, lib.rs:56-67)
    let pts = vec![g, g + g, -g, G::zero(), g * F::from(5u64), -(g * F::from(5u64))];
    j.push(format!(
        "\"points\":{{\"multiples_of_g\":[1,2,-1,0,5,-5],\"compressed\":{},\"xy\":{},\"to_data_item\":{}}}",
        list(pts.iter().map(pc).collect()),
        list(pts.iter().map(pu).collect()),
        list(pts.iter().map(|p| fr(&p.to_data_item())).collect())
    ));

    // 2. DefaultFieldHasher (transcript.rs:55): hash_to_field(msg, 1)[0] under the labels the crate uses
    let mut h = Vec::new();
    for (dst, msg) in [("ipa", "abc"), ("multiproof", ""), ("ipa", "0123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef0123456789")] {
        let hasher = <Hasher as HashToField<F>>::new(dst.as_bytes());
        let v: F = hasher.hash_to_field(msg.as_bytes(), 1)[0];
        h.push(format!("{{\"dst\":\"{}\",\"msg\":\"{}\",\"out\":{}}}", dst, msg, fr(&v)));
    }
    j.push(format!("\"hash_to_field\":{}", list(h)));

    // 3. TranscriptHasher walk of an IPA opening (ipa/mod.rs:286-306)
    {
        let (c, z, y, l, r) = (g * F::from(5u64), F::from(7u64), F::from(11u64), g + g, G::zero());
        let mut t = T::new("ipa");
        t.append(&c, "C").unwrap();
        t.append(&z, "input point").unwrap();
        t.append(&y, "output point").unwrap();
        let w = t.digest("w", true);
        t.append(&l, "L").unwrap();
        t.append(&r, "R").unwrap();
        let x = t.digest("x", true);
        t.append(&1234567usize, "z").unwrap(); // usize as multiproof.rs:111 appends it
        let u = t.digest("u", false);
        let u2 = t.digest("u", true);
        j.push(format!(
            "\"transcript\":{{\"c\":5,\"z\":7,\"y\":11,\"l\":2,\"r\":0,\"w\":{},\"x\":{},\"usize\":1234567,\"u_noclear\":{},\"u_clear\":{}}}",
            fr(&w),
            fr(&x),
            fr(&u),
            fr(&u2)
        ));
    }

    // 4. radix-2 domain generators (precompute.rs:26-27)
    let mut d = Vec::new();
    for n in [2usize, 4, 16, 32, 256, 1 << 14] {
        d.push(format!("{{\"n\":{},\"group_gen\":{}}}", n, fr(&Dom::new(n).unwrap().group_gen())));
    }
    j.push(format!("\"domains\":{}", list(d)));

    // 5. IPA CRS (ipa_point_generator.rs:51-109): the first points of the default generator
    {
        let crs: Vec<G> = IPAPointGenerator::<G, crate::ipa::ipa_point_generator::EthereumHashToCurve>::default().gen(8).unwrap();
        j.push(format!("\"ipa_crs\":{{\"seed\":\"eth_verkle_oct_2021\",\"xy\":{}}}", list(crs.iter().map(pu).collect())));
    }

    // 6. IPA: commit, prove (inside / outside the domain), prove_commitment
    let mut cases = Vec::new();
    ipa_case::<4>(&mut cases);
    ipa_case::<32>(&mut cases);
    j.push(format!("\"ipa\":{}", list(cases)));

    // 7. KZG (kzg/mod.rs:278-297 shape): tau = 100, key 16, data 8 over the key's domain
    {
        type K = KZG<Bn254, Hasher, Dom>;
        let key = K::setup(16, &KZGRandomPointGenerator::<G>::default()).unwrap();
        let dt = LagrangeBasis::<F, Dom>::from_vec_and_domain(data(8, 9, 5), Dom::new(16).unwrap());
        let c = K::commit(&key, &dt).unwrap();
        let mut pr = Vec::new();
        for idx in [0usize, 3, 7, 9, 17, 40] {
            let p = K::prove(&key, &c, idx, &dt).unwrap();
            assert!(K::verify(&key, &c, idx, &p).unwrap());
            pr.push(format!("{{\"index\":{},\"proof\":{},\"y\":{}}}", idx, pc(&p.proof), fr(&p.y)));
        }
        j.push(format!(
            "\"kzg\":{{\"tau\":100,\"key\":16,\"data\":\"9+5i, 8 values\",\"lagrange_xy\":{},\"commit\":{},\"proofs\":{}}}",
            list(key.lagrange_commitments.iter().map(pu).collect()),
            pc(&c),
            list(pr)
        ));
    }

    // 8. IPA multiproof (multiproof.rs:261-308 shape): 5 vectors of width 32, two of them opened at the same point
    {
        type I = IPA<32, G, Hasher, Dom>;
        let key = IPAUniversalParams::<32, G, Hasher>::new_from_vec(gens(32));
        let all: Vec<(LagrangeBasis<F, Dom>, G)> = (0..5u64)
            .map(|q| {
                let dt = LagrangeBasis::<F, Dom>::from_vec(data(32, 100 * q + 1, q + 2));
                let c = I::commit(&key, &dt).unwrap();
                (dt, c)
            })
            .collect();
        let zs = [3usize, 31, 3, 0, 17];
        let queries: Vec<_> = all
            .iter()
            .zip(zs.iter())
            .map(|((dt, c), z)| MultiproofProverQuery::new(dt, c, *z, dt[*z]))
            .collect();
        let vq: Vec<_> = queries.iter().map(|q| q.to_verifier_query()).collect();
        let mp = I::prove_multiproof(&key, &queries).unwrap();
        assert!(I::verify_multiproof(&key, &vq, &mp).unwrap());
        j.push(format!(
            "\"multiproof\":{{\"n\":32,\"data\":\"(100q+1)+(q+2)i\",\"z\":[3,31,3,0,17],\"commits\":{},\"d\":{},\"l\":{},\"r\":{},\"tip\":{},\"y\":{}}}",
            list(all.iter().map(|(_, c)| pc(c)).collect()),
            pc(&mp.d),
            list(mp.proof.l.iter().map(pc).collect()),
            list(mp.proof.r.iter().map(pc).collect()),
            fr(&mp.proof.tip),
            fr(&mp.proof.y)
        ));
    }

    println!("ARKVEC_BEGIN");
    println!("{{{}}}", j.join(","));
    println!("ARKVEC_END");
    let _ = F::one();
}
