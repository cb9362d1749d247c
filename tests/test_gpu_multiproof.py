"""P1 / P2 parity on the GPU: multiproof.rs:261-357 restated (prove, verify, tamper) + oracle byte equality."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from verkle_kzg_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _queries(eng, key, rng, N, m, zs=None):
    f = orc.rand_fr_buf(rng, m * N).reshape(m, N, 32)
    C = eng.commit_batch(key, f)
    z = np.array(zs if zs is not None else rng.integers(0, N, m), dtype=np.uint64)
    y = np.stack([f[i, int(z[i])] for i in range(m)])
    return f, C, z, y


@pytest.mark.parametrize("N,m,wb", [(32, 20, 8), (32, 1, 8), (256, 64, 0)])
def test_ipa_multiproof(eng, N, m, wb):
    rng = np.random.default_rng(500 + N + m)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N], window_bits=wb)
    f, C, z, y = _queries(eng, key, rng, N, m)
    got = eng.multiproof_prove(key, "ipa", f, C, z, y)
    exp = orc.multiproof_prove("ipa", bases, N, f, C, z, y)
    for k in ("D", "L", "R", "tip", "y"):
        assert (got[k] == exp[k]).all(), k
    assert orc.multiproof_verify("ipa", bases, N, C, z, y, got)
    assert eng.multiproof_verify_ipa(key, C, z, y, got)
    # tamper d and y (multiproof.rs:296-307)
    bad = dict(got)
    bad["D"] = bases[0]
    assert not eng.multiproof_verify_ipa(key, C, z, y, bad)
    y2 = y.copy()
    y2[0] = orc.field_op(0, "add", y[0], orc.fr_to_buf([1])[0])[0]
    assert not eng.multiproof_verify_ipa(key, C, z, y2, got)
    key.free()


def test_ipa_multiproof_skewed_groups(eng):
    """all queries at one z (a single long group -> several segments), duplicates of one commitment, and z = N - 1"""
    N, m = 32, 100
    rng = np.random.default_rng(77)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N], window_bits=8)
    f, C, z, y = _queries(eng, key, rng, N, m, zs=[5] * 70 + [N - 1] * 30)
    f[1] = f[0]
    C[1] = C[0]
    y[1] = y[0]
    got = eng.multiproof_prove(key, "ipa", f, C, z, y)
    exp = orc.multiproof_prove("ipa", bases, N, f, C, z, y)
    for k in ("D", "L", "R", "tip", "y"):
        assert (got[k] == exp[k]).all(), k
    assert eng.multiproof_verify_ipa(key, C, z, y, got)
    key.free()


def test_ipa_multiproof_many_queries(eng):
    """more commitments than the four-lane scalar multiplication takes (> 64 per SM): the thread-per-point windowed form"""
    N, m = 32, 12000
    rng = np.random.default_rng(79)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N], window_bits=8)
    f, C, z, y = _queries(eng, key, rng, N, m)
    got = eng.multiproof_prove(key, "ipa", f, C, z, y)
    assert eng.multiproof_verify_ipa(key, C, z, y, got)
    assert orc.multiproof_verify("ipa", bases, N, C, z, y, got)
    y2 = y.copy()
    y2[m - 1] = orc.field_op(0, "add", y[m - 1], orc.fr_to_buf([1])[0])[0]
    assert not eng.multiproof_verify_ipa(key, C, z, y2, got)
    key.free()


def test_kzg_multiproof(eng):
    N, m = 32, 20
    rng = np.random.default_rng(78)
    srs = orc.kzg_setup(N, 100)
    key = eng.load_key(srs, window_bits=8)
    f, C, z, y = _queries(eng, key, rng, N, m)
    got = eng.multiproof_prove(key, "kzg", f, C, z, y)
    exp = orc.multiproof_prove("kzg", srs, N, f, C, z, y)
    assert (got["D"] == exp["D"]).all() and (got["L"][0] == exp["L"][0]).all() and (got["y"] == exp["y"]).all()
    assert orc.multiproof_verify("kzg", srs, N, C, z, y, got, tau=100)
    key.free()


def test_multiproof_rejects_out_of_range_z(eng):
    from verkle_kzg_b200 import VkzgError
    N = 32
    rng = np.random.default_rng(79)
    srs = orc.kzg_setup(N, 100)
    key = eng.load_key(srs, window_bits=8)
    f, C, z, y = _queries(eng, key, rng, N, 3)
    z[1] = N
    with pytest.raises(VkzgError):
        eng.multiproof_prove(key, "kzg", f, C, z, y)
    key.free()


def test_trait_surface_ipa(eng):
    from verkle_kzg_b200.vector_commit import IPA, LagrangeBasis
    N = 32
    rng = np.random.default_rng(80)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = IPA.setup(eng, bases[:N], q=bases[N], window_bits=8)
    data = LagrangeBasis.from_vec(orc.rand_fr_buf(rng, N))
    C = IPA.commit(key, data)
    idx = 7
    pf = IPA.prove(key, C, idx, data)
    assert IPA.verify(key, C, idx, pf)
    out = IPA.prove(key, C, 2 * N, data)            # outside the domain (ipa/mod.rs:413-416)
    assert IPA.verify(key, C, 2 * N, out)
    assert not IPA.verify(key, C, idx, out)
    cp = IPA.prove_commitment(key, C, data)
    assert orc.ipa_verify_commitment(bases, N, C, cp["l"], cp["r"], cp["tip"])
    queries = []
    for i in range(20):
        d = LagrangeBasis.from_vec(orc.rand_fr_buf(rng, N))
        z = int(rng.integers(0, N))
        queries.append((d, IPA.commit(key, d), z, d.evaluations[z]))
    mp = IPA.prove_multiproof(key, queries)
    vq = [(q[1], q[2], q[3]) for q in queries]
    assert IPA.verify_multiproof(key, vq, mp)
    key.free()


def test_claimed_evaluations_are_not_bound_by_the_reference_protocol(eng):
    """the reference's verify_multiproof computes g2(t) = sum r^q y_q / (t - z_q) and never compares it with the
    proof's evaluation (multiproof.rs:201-215), so a prover may claim ANY y_q — reproduced (same verdict as the oracle).
    The comparison cannot simply be switched on: the reference's prover divides by (X - w^z) in g and by (t - z), z an
    integer, in h (quirk Q4), so (h - g)(t) != g2(t) for HONEST proofs as well; the diagnostic option shows exactly that."""
    rng = np.random.default_rng(99)
    N, m = 32, 9
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N], window_bits=12)
    f = orc.rand_fr_buf(rng, m * N).reshape(m, N, 32)
    C = eng.commit_batch(key, f)
    z = rng.integers(2, N, m).astype(np.uint64)
    y = f[np.arange(m), z.astype(np.int64)]
    good = eng.multiproof_prove(key, "ipa", f, C, z, y)
    assert eng.multiproof_verify_ipa(key, C, z, y, good)
    lie = y.copy()
    lie[4] = orc.field_op(0, "add", y[4], orc.fr_to_buf([1])[0])[0]       # a wrong claimed evaluation, consistently in the transcript
    forged = eng.multiproof_prove(key, "ipa", f, C, z, lie)
    assert orc.multiproof_verify("ipa", bases, N, C, z, lie, forged)       # the reference's algorithm accepts it ...
    assert eng.multiproof_verify_ipa(key, C, z, lie, forged)               # ... and so does the drop-in
    eng.set_option(eng.OPT_MULTIPROOF_CHECK_Y, 1)
    try:
        assert not eng.multiproof_verify_ipa(key, C, z, lie, forged)
        assert not eng.multiproof_verify_ipa(key, C, z, y, good)           # honest proofs fail the comparison too (Q4)
    finally:
        eng.set_option(eng.OPT_MULTIPROOF_CHECK_Y, 0)
    assert eng.multiproof_verify_ipa(key, C, z, y, good)
    key.free()


@pytest.mark.parametrize("scheme", ["ipa", "kzg"])
def test_batched_multiproofs_equal_single_calls(eng, scheme):
    """vkzg_multiproof_prove_batch: K multiproofs of different sizes (1 query, skewed points, > 8 proofs so side streams are
    reused) are byte-identical to K calls of vkzg_multiproof_prove and to the oracle; IPA ones verify on the device"""
    rng = np.random.default_rng(321)
    N = 32
    if scheme == "ipa":
        k0, k1 = orc.rand_fr(rng, 2)
        bases = orc.points_walk(k0, k1, N + 1)
        key = eng.load_key(bases[:N], q=bases[N], window_bits=10)
    else:
        bases = orc.kzg_setup(N, 100)
        key = eng.load_key(bases, window_bits=10)
    m_each = [5, 1, 40, 17, 3, 64, 2, 9, 33, 12, 7]
    fs, Cs, zs, ys = [], [], [], []
    for i, m in enumerate(m_each):
        f, C, z, y = _queries(eng, key, rng, N, m, zs=None if i != 2 else [3] * 30 + [7] * 10)
        fs.append(f), Cs.append(C), zs.append(z), ys.append(y)
    got = eng.multiproof_prove_batch(key, scheme, np.concatenate(fs), np.concatenate(Cs), np.concatenate(zs), np.concatenate(ys), m_each)
    assert len(got) == len(m_each)
    for i in range(len(m_each)):
        one = eng.multiproof_prove(key, scheme, fs[i], Cs[i], zs[i], ys[i])
        assert all((got[i][k] == one[k]).all() for k in one), f"proof {i} differs from the single call"
        if i in (0, 2, 5):
            exp = orc.multiproof_prove(scheme, bases, N, fs[i], Cs[i], zs[i], ys[i])
            assert all((got[i][k] == exp[k]).all() for k in exp), f"proof {i} differs from the oracle"
        if scheme == "ipa":
            assert eng.multiproof_verify_ipa(key, Cs[i], zs[i], ys[i], got[i])
    key.free()
