"""SURVEY 8f-4 on the GPU: IPAPointGenerator::gen / gen_at with EthereumHashToCurve (ipa/ipa_point_generator.rs:51-109)
through vkzg_ipa_crs_generate[_at], byte for byte against the oracle and the Python model."""
import numpy as np
import pytest

import orc
import pyref

pytestmark = pytest.mark.gpu

DEFAULT_SEED = b"eth_verkle_oct_2021"  # IPAPointGenerator::default (ipa_point_generator.rs:38-47)


@pytest.fixture(scope="module")
def eng():
    from verkle_kzg_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("seed,num", [(DEFAULT_SEED, 256), (DEFAULT_SEED, 1), (b"", 33), (bytes(range(200)), 100), (b"s" * 64, 17),
                                      (b"t" * 55, 5), (b"u" * 56, 5), (b"v" * 119, 5)])
def test_gen_matches_oracle(eng, seed, num):
    got, nxt = eng.ipa_crs_generate(seed, num)
    want, want_next = orc.ipa_crs_gen(seed, num)
    assert nxt == want_next
    assert (got == want).all()


def test_gen_matches_python_model_and_gen_at(eng):
    want, nxt = pyref.ipa_crs_gen(DEFAULT_SEED, 24)
    got, got_next = eng.ipa_crs_generate(DEFAULT_SEED, 24)
    assert got_next == nxt and (got == orc.pts_to_buf(want)).all()
    hits = []
    for i in range(nxt):
        p = eng.ipa_crs_generate_at(DEFAULT_SEED, i)
        ref = orc.ipa_crs_gen_at(DEFAULT_SEED, i)
        assert (p is None) == (ref is None)
        if p is not None:
            assert (p == ref).all()
            hits.append(p)
    assert len(hits) == 24 and all((a == b).all() for a, b in zip(hits, got))


def test_large_crs_is_a_prefix_chain_and_usable_as_a_key(eng):
    """several candidate passes (the request exceeds one pass), prefix property, and the points work as an IPA key"""
    big, nxt_big = eng.ipa_crs_generate(b"big", 20000)
    small, nxt_small = eng.ipa_crs_generate(b"big", 3000)
    assert (big[:3000] == small).all() and nxt_small < nxt_big
    want, want_next = orc.ipa_crs_gen(b"big", 3000)
    assert (small == want).all() and nxt_small == want_next
    assert len({bytes(r) for r in big}) == len(big)
    N = 32
    key = eng.load_key(big[:N], q=big[N], window_bits=8)
    rng = np.random.default_rng(3)
    a = orc.rand_fr_buf(rng, N).reshape(1, N, 32)
    C = eng.commit_batch(key, a)
    assert (C == orc.commit_batch(big[:N], a)).all()
    z = orc.fr_to_buf([5])
    L, R, tip, y = eng.ipa_prove_batch(key, a, z, C)
    assert eng.ipa_verify_batch(key, z, C, L, R, tip, y).all()
    key.free()


def test_point_generator_mirror(eng):
    """the reference's interface: max / OutOfBounds / InvalidPoint / secret (ipa_point_generator.rs:21-86)"""
    from verkle_kzg_b200.vector_commit import IPAPointGenerator, OutOfBounds, InvalidPoint
    g = IPAPointGenerator(eng)
    assert g.secret() == DEFAULT_SEED and g.max == 256
    pts = g.gen(8)
    assert (pts == orc.ipa_crs_gen(DEFAULT_SEED, 8)[0]).all()
    with pytest.raises(OutOfBounds):
        g.gen(257)
    g.set_max(300)
    assert len(g.gen(257)) == 257
    with pytest.raises(OutOfBounds):
        g.gen_at(301)
    first_bad = next(i for i in range(50) if orc.ipa_crs_gen_at(DEFAULT_SEED, i) is None)
    with pytest.raises(InvalidPoint):
        g.gen_at(first_bad)
    first_ok = next(i for i in range(50) if orc.ipa_crs_gen_at(DEFAULT_SEED, i) is not None)
    assert (g.gen_at(first_ok) == pts[0]).all()


def test_errors(eng):
    from verkle_kzg_b200._lib import VkzgError
    with pytest.raises(VkzgError):
        eng.ipa_crs_generate(b"x", 0)


def test_setup_from_generators_like_the_reference(eng):
    """IPA::setup(max_items, &IPAPointGenerator) (ipa/mod.rs:121-128) and KZG::setup(max_items, &KZGRandomPointGenerator)
    (kzg/mod.rs:115-124) through the mirror: keys made on the GPU prove and verify like keys made from oracle points"""
    from verkle_kzg_b200.vector_commit import IPA, KZG, IPAPointGenerator, KZGRandomPointGenerator, LagrangeBasis, OutOfBounds
    N = 32
    gen = IPAPointGenerator(eng, max=N + 1)
    key = IPA.setup_from_generator(eng, N, gen, window_bits=8)
    bases = orc.ipa_crs_gen(DEFAULT_SEED, N + 1)[0]
    rng = np.random.default_rng(11)
    data = LagrangeBasis(orc.rand_fr_buf(rng, N))
    C = IPA.commit(key, data)
    assert (C == orc.commit_batch(bases[:N], data.evaluations[None])[0]).all()
    proof = IPA.prove(key, C, 3, data)
    eL, eR, etip, ey = orc.ipa_prove(bases, N, data.evaluations, C, orc.fr_to_buf([3])[0])
    assert (proof["l"] == eL).all() and (proof["r"] == eR).all() and (proof["tip"] == etip).all() and (proof["y"] == ey).all()
    assert IPA.verify(key, C, 3, proof)
    with pytest.raises(OutOfBounds):
        IPA.setup_from_generator(eng, N + 1, gen)
    key.free()
    kgen = KZGRandomPointGenerator(eng)            # secret 100, kzg_point_generator.rs:20-26
    assert kgen.secret() == 100
    assert (kgen.gen(5) == orc.g1_mul_gen_batch(orc.fr_to_buf([pow(100, i, orc.R_MOD) for i in range(5)]))).all()
    kkey = KZG.setup_from_generator(eng, 16, kgen, window_bits=8)
    srs = orc.kzg_setup(16, 100)
    d8 = LagrangeBasis.from_vec_and_domain(orc.rand_fr_buf(rng, 8), 16)
    Ck = KZG.commit(kkey, d8)
    assert (Ck == orc.commit_batch(srs[:8], d8.evaluations[None])[0]).all()
    pf = KZG.prove(kkey, Ck, 5, d8)
    epf, ey, ok = orc.kzg_prove(srs, d8.evaluations, orc.fr_to_buf([5])[0])
    assert ok and (pf["proof"] == epf).all() and (pf["y"] == ey).all()
    kkey.free()
