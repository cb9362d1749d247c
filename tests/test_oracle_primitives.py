"""Oracle pinning, part 1: primitives against known answers and the independent Python model."""
import hashlib

import numpy as np
import pytest

import orc
import pyref

RFC9380_DST = b"QUUX-V01-CS02-with-expander-SHA256-128"


def test_sha256_fips_vectors():
    assert orc.sha256(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert orc.sha256(b"").hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    m = b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"
    assert orc.sha256(m).hex() == "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"
    rng = np.random.default_rng(1)
    for n in [1, 55, 56, 63, 64, 65, 119, 120, 121, 127, 128, 129, 1000]:
        msg = rng.bytes(n)
        assert orc.sha256(msg) == hashlib.sha256(msg).digest()


@pytest.mark.parametrize("msg,want", [
    (b"", "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"),
    (b"abc", "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"),
    (b"abcdef0123456789", "eff31487c770a893cfb36f912fbfcbff40d5661771ca4b2cb4eafe524333f5c1"),
])
def test_xmd_rfc9380_vectors(msg, want):
    # RFC 9380 appendix K.1 (z_pad = SHA-256 block size 64)
    assert orc.expand_message_xmd(msg, RFC9380_DST, 32, 64).hex() == want
    assert pyref.expand_message_xmd(msg, RFC9380_DST, 32, 64).hex() == want


def test_xmd_ark04_variant_matches_python_model():
    rng = np.random.default_rng(2)
    for n in [0, 1, 33, 100, 121, 187, 300]:
        msg = rng.bytes(n)
        for dst in [b"ipa", b"multiproof"]:
            assert orc.expand_message_xmd(msg, dst, 48, 48) == pyref.expand_message_xmd(msg, dst, 48, 48)
            got = orc.buf_to_fr(orc.hash_to_fr(msg, dst.decode()))[0]
            assert got == pyref.hash_to_fr(msg, dst)


def test_field_ops_match_python_ints():
    rng = np.random.default_rng(3)
    for tag, mod, enc, dec in [(0, orc.R_MOD, orc.fr_to_buf, orc.buf_to_fr), (1, orc.P_MOD, orc.fq_to_buf, orc.buf_to_fq)]:
        xs = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(64)] + [0, 1, mod - 1, mod - 2]
        ys = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(64)] + [mod - 1, 0, mod - 1, 2]
        a, b = enc(xs), enc(ys)
        assert dec(orc.field_op(tag, "add", a, b)) == [(x + y) % mod for x, y in zip(xs, ys)]
        assert dec(orc.field_op(tag, "sub", a, b)) == [(x - y) % mod for x, y in zip(xs, ys)]
        assert dec(orc.field_op(tag, "mul", a, b)) == [(x * y) % mod for x, y in zip(xs, ys)]
        nz = [x for x in xs if x]
        assert dec(orc.field_op(tag, "inv", enc(nz))) == [pow(x, -1, mod) for x in nz]
        # Montgomery conversion round trip against the pure-python encoder
        canon = np.stack([np.frombuffer(x.to_bytes(32, "little"), dtype=np.uint8) for x in xs])
        assert (orc.to_mont(tag, canon) == a).all()
        assert (orc.from_mont(tag, a) == canon).all()


def test_montgomery_constants_from_survey():
    # R mod r and R mod p
    assert orc.MONT_R % orc.R_MOD == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert orc.MONT_R % orc.P_MOD == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d
    one = orc.fr_to_buf([1])[0]
    assert int.from_bytes(bytes(one), "little") == orc.MONT_R % orc.R_MOD


def test_g1_known_multiples():
    # alt_bn128 2G (EIP-196 test vectors)
    g = orc.g1_generator()
    assert orc.buf_to_pts(g) == [(1, 2)]
    two_g = orc.buf_to_pts(orc.g1_add(g, g))[0]
    assert two_g == (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
                     0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)
    assert two_g == pyref.g_add(pyref.G1_GEN, pyref.G1_GEN)
    # r * G = identity, (r-1) G = -G
    assert orc.buf_to_pts(orc.g1_mul(g, orc.fr_to_buf([orc.R_MOD - 1])))[0] == (1, orc.P_MOD - 2)
    assert orc.buf_to_pts(orc.g1_add(orc.g1_mul(g, orc.fr_to_buf([orc.R_MOD - 1])), g))[0] is None


def test_g1_group_law_matches_python_model():
    rng = np.random.default_rng(4)
    ks = orc.rand_fr(rng, 12) + [0, 1, 2, orc.R_MOD - 1]
    pts = [pyref.g_mul(pyref.G1_GEN, k) for k in ks]
    bufs = orc.pts_to_buf(pts)
    got = orc.buf_to_pts(orc.g1_mul_gen_batch(orc.fr_to_buf(ks), 2))
    assert got == pts
    for i in range(len(pts)):
        for j in [0, 3, len(pts) - 1, i]:
            assert orc.buf_to_pts(orc.g1_add(bufs[i], bufs[j]))[0] == pyref.g_add(pts[i], pts[j])
        assert orc.g1_on_curve(bufs[i])
    walk = orc.buf_to_pts(orc.points_walk(ks[0], ks[1], 5))
    assert walk == [pyref.g_mul(pyref.G1_GEN, ks[0] + i * ks[1]) for i in range(5)]


def test_compressed_serialisation_and_to_data_item():
    rng = np.random.default_rng(5)
    ks = orc.rand_fr(rng, 16)
    pts = [pyref.g_mul(pyref.G1_GEN, k) for k in ks] + [None, pyref.G1_GEN, pyref.g_neg(pyref.G1_GEN)]
    buf = orc.pts_to_buf(pts)
    comp = orc.g1_compress(buf)
    for row, p in zip(comp, pts):
        assert bytes(row) == pyref.ser_g1(p)
    # generator: y = 2 <= p - 2  -> no flag ; -G -> 0x80 ; infinity -> 0x40
    assert bytes(comp[-2]) == (1).to_bytes(32, "little")
    assert comp[-1][31] == 0x80 and comp[-3][31] == 0x40
    assert orc.buf_to_fr(orc.to_data_item(buf)) == [pyref.to_data_item(p) for p in pts]


def test_domain_generators():
    # SURVEY.md section 8 constants
    assert orc.buf_to_fr(orc.domain_gen(256))[0] == 3478517300119284901893091970156912948790432420133812234316178878452092729974
    assert orc.buf_to_fr(orc.domain_gen(32))[0] == 4419234939496763621076330863786513495701855246241724391626358375488475697872
    assert pyref.group_gen(256) == 3478517300119284901893091970156912948790432420133812234316178878452092729974
    w = pyref.group_gen(256)
    assert pow(w, 256, pyref.R_MOD) == 1 and pow(w, 128, pyref.R_MOD) != 1
    assert orc.buf_to_fr(orc.domain_gen(20))[0] == pyref.group_gen(32)


# Public alt_bn128 vectors (EIP-196 precompiles 0x06 / 0x07; go-ethereum's bn256Add / bn256ScalarMul test files,
# case "chfast1") — produced by independent implementations, not by this repository's oracle or Python model.
EIP196_ADD = (("18b18acfb4c2c30276db5411368e7185b311dd124691610c5d3b74034e093dc9", "063c909c4720840cb5134cb9f59fa749755796819658d32efc0d288198f37266"),
              ("07c2b7f58a84bd6145f00c9c2bc0bb1a187f20ff2c92963a88019e7c6a014eed", "06614e20c147e940f2d70da3f74c9a17df361706a4485c742bd6788478fa17d7"),
              ("2243525c5efd4b9c3d3c45ac0ca3fe4dd85e830a4ce6b65fa1eeaee202839703", "301d1d33be6da8e509df21cc35964723180eed7532537db9ae5e7d48f195c915"))
EIP196_MUL = (("2bd3e6d0f3b142924f5ca7b49ce5b9d54c4703d7ae5648e61d02268b1a0a9fb7", "21611ce0a6af85915e2f1d70300909ce2e49dfad4a4619c8390cae66cefdb204"),
              0x11138ce750fa15c2,
              ("070a8d6a982153cae4be29d434e8faef8a47b274a053f5a4ee2a6c9c13c31e5c", "031b8ce914eba3a9ffb989f9cdd5b0f01943074bf4f0f315690ec3cec6981afc"))
EIP196_2G = ("030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3", "15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4")


def eip196_point(xy):
    return orc.pts_to_buf([(int(xy[0], 16), int(xy[1], 16))])[0]


def test_group_law_against_public_eip196_vectors():
    a, b, s = (eip196_point(p) for p in EIP196_ADD)
    assert orc.g1_on_curve(a) and orc.g1_on_curve(b)
    assert (orc.g1_add(a, b) == s).all() and (orc.g1_add(b, a) == s).all()
    p, k, r = eip196_point(EIP196_MUL[0]), EIP196_MUL[1], eip196_point(EIP196_MUL[2])
    assert (orc.g1_mul(p, orc.fr_to_buf([k])[0]) == r).all()
    assert (orc.msm(p[None], orc.fr_to_buf([k]), mode="naive") == r).all()
    assert (orc.msm(p[None], orc.fr_to_buf([k]), mode="pippenger") == r).all()
    g = orc.g1_generator()
    assert (orc.g1_add(g, g) == eip196_point(EIP196_2G)).all()
    assert (orc.g1_mul(g, orc.fr_to_buf([2])[0]) == eip196_point(EIP196_2G)).all()


# Public ark-bn254 0.4 compressed-point vectors: the test constants of aptos-core's
# aptos-move/framework/aptos-stdlib/sources/cryptography/bn254_algebra.move (format `FormatG1Compr`, documented there as
# "x in little-endian; if y > -y set bit 0b1000_0000 of the last byte; infinity sets 0b0100_0000" and implemented with
# ark-bn254 `serialize_compressed`).  They are typed in from memory of that file (no network here): the coordinates are
# checked against the group law below, so only the ASSIGNMENT of the flag to 7G / -7G rests on recollection — evidence
# for the 0x80 <=> y > -y convention of g1_affine_serialize_compressed / affine_compress, not a substitute for the
# tests/golden/arkworks kit.
APTOS_G1_GENERATOR_COMP = "0100000000000000000000000000000000000000000000000000000000000000"
APTOS_G1_INF_COMP = "0000000000000000000000000000000000000000000000000000000000000040"
APTOS_G1_7G_UNCOMP = ("78e0ffab866b3a9876bd01b8ecc66fcb86936277f425539a758dbbd32e2b0717"
                      "9eafd4607f9f80771bf4185df03bfead7a3719fa4bb57b0152dd30d16cda8a16")
APTOS_G1_7G_COMP = "78e0ffab866b3a9876bd01b8ecc66fcb86936277f425539a758dbbd32e2b0717"
APTOS_G1_NEG_7G_COMP = "78e0ffab866b3a9876bd01b8ecc66fcb86936277f425539a758dbbd32e2b0797"


def test_compressed_flags_against_public_ark_bn254_vectors():
    g7 = pyref.g_mul(pyref.G1_GEN, 7)
    raw = bytes.fromhex(APTOS_G1_7G_UNCOMP)
    assert (int.from_bytes(raw[:32], "little"), int.from_bytes(raw[32:], "little")) == g7   # the recalled coordinates are 7G
    buf = orc.pts_to_buf([pyref.G1_GEN, None, g7, pyref.g_neg(g7)])
    comp = [bytes(r).hex() for r in orc.g1_compress(buf)]
    assert comp == [APTOS_G1_GENERATOR_COMP, APTOS_G1_INF_COMP, APTOS_G1_7G_COMP, APTOS_G1_NEG_7G_COMP]
    assert [pyref.ser_g1(p).hex() for p in (pyref.G1_GEN, None, g7, pyref.g_neg(g7))] == comp
    # from_random_bytes is the inverse of the encoder on these
    for p, h in ((pyref.G1_GEN, APTOS_G1_GENERATOR_COMP), (g7, APTOS_G1_7G_COMP), (pyref.g_neg(g7), APTOS_G1_NEG_7G_COMP)):
        assert pyref.from_random_bytes(bytes.fromhex(h)) == p
    assert pyref.from_random_bytes(bytes.fromhex(APTOS_G1_INF_COMP)) == "inf"


def test_ipa_crs_generation_matches_python_model():
    """IPAPointGenerator::gen / gen_at (ipa_point_generator.rs:51-81) with EthereumHashToCurve (:97-109)"""
    for seed, num in ((b"eth_verkle_oct_2021", 40), (b"", 5), (bytes(range(70)), 9), (b"x" * 64, 7)):
        want, nxt = pyref.ipa_crs_gen(seed, num)
        got, got_next = orc.ipa_crs_gen(seed, num)
        assert got_next == nxt
        assert (got == orc.pts_to_buf(want)).all()
        assert all(orc.g1_on_curve(p) for p in got)
        # gen_at agrees index by index, and gen is exactly the valid indices in order
        hits = []
        for i in range(nxt):
            p = orc.ipa_crs_gen_at(seed, i)
            if p is not None:
                hits.append(p)
        assert len(hits) == num and all((a == b).all() for a, b in zip(hits, got))
    # acceptance rate ~ 1/2 (flags) x p / 2^254 x 1/2 (quadratic residue) = 0.189
    _, nxt = orc.ipa_crs_gen(b"eth_verkle_oct_2021", 256)
    assert 1000 < nxt < 1800
    # every output is a fixed point of serialise -> from_random_bytes (the flag convention is shared)
    pts, _ = orc.ipa_crs_gen(b"round-trip", 12)
    for row, comp in zip(pts, orc.g1_compress(pts)):
        back = pyref.from_random_bytes(bytes(comp))
        assert (orc.pts_to_buf([back])[0] == row).all()
