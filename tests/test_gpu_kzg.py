"""E1 / K1 / K2 / K3 parity on the GPU (kzg/mod.rs:278-297 restated + oracle byte equality)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu

TAU = 100  # kzg_point_generator.rs:20-26


@pytest.fixture(scope="module")
def eng():
    from verkle_kzg_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _key(eng, n, wb):
    srs = orc.kzg_setup(n, TAU)
    return srs, eng.load_key(srs, window_bits=wb)


@pytest.mark.parametrize("N,length,domain,wb", [(16, 8, 16, 8), (16, 8, 0, 8), (16, 16, 0, 8), (32, 20, 32, 8), (32, 20, 0, 8),
                                                 (256, 256, 0, 0)])
def test_quotient_evaluate_open_match_oracle(eng, N, length, domain, wb):
    """the reference's own test shape (DATA_SIZE 8 over the key's domain of 16, kzg/mod.rs:266-274), from_vec's own
    domain, its bench shape (20 in 32) and width 256"""
    srs, key = _key(eng, N, wb)
    rng = np.random.default_rng(N * 1000 + length + domain)
    want = domain if domain else length
    dn = 1
    while dn < want:
        dn <<= 1
    pts = list(range(0, min(length, 6))) + [length - 1]
    if dn > length:
        pts += [length, dn - 1]          # stored range exceeded but inside the data domain -> y = 0
    if dn < N:
        pts += [dn + 1, N - 1]           # beyond the data domain, below the key size
    pts += [N + 1, 2 * N, orc.R_MOD - 2] + orc.rand_fr(rng, 2)  # (r - 1 = w^(n/2) is a root of unity: the reference divides by zero)
    pts = [p for p in pts if not (p <= N and p >= dn)]  # the reference panics there (checked in the next test)
    B = len(pts)
    f = orc.rand_fr_buf(rng, B * length).reshape(B, length, 32)
    zb = orc.fr_to_buf(pts)
    ys = eng.evaluate_batch(key, f, zb, domain_n=domain)
    q, y2 = eng.quotient_batch(key, f, zb, domain_n=domain)
    proof, y3 = eng.kzg_open_batch(key, f, zb, domain_n=domain)
    C = eng.commit_batch(key, f)
    for i, p in enumerate(pts):
        ey = orc.evaluate(N, f[i], want, zb[i])
        assert (ys[i] == ey).all() and (y2[i] == ey).all() and (y3[i] == ey).all(), p
        if p <= N:
            eq = orc.divide_by_vanishing(N, f[i], want, p)
        else:
            eq = orc.divide_by_vanishing_outside(N, f[i], want, zb[i])
        assert (q[i] == eq).all(), p
        # proof = <SRS, q> (kzg/mod.rs:148-149)
        assert (proof[i] == orc.msm(srs, eq)).all(), p
        if dn == N:
            epf, ey2, ok = orc.kzg_prove(srs, f[i], zb[i])  # the oracle's prove_point uses the key's domain
            assert ok and (proof[i] == epf).all() and (y3[i] == ey2).all(), p
            # pairing-free check with the known tau:  [tau - z] pi == C - [y] G   (oracle, kzg/mod.rs:165-189)
            assert orc.kzg_verify_tau(srs, TAU, C[i], zb[i], proof[i], y3[i]), p
    key.free()


def test_commit_open_fused_equals_separate_calls(eng):
    """vkzg_kzg_commit_open_batch (one upload of the rows) == vkzg_commit_batch + vkzg_kzg_open_batch, incl. the chunked path"""
    srs, key = _key(eng, 32, 8)
    rng = np.random.default_rng(4321)
    for B, length in ((1, 32), (37, 20), (9000, 32)):   # 9000 >= 8192: the rows upload in pipelined chunks
        f = orc.rand_fr_buf(rng, B * length).reshape(B, length, 32)
        zb = orc.fr_to_buf([int(v) for v in rng.integers(0, length, B)])
        zb[B // 2] = orc.fr_to_buf([77777])[0]               # one outside point
        C, pf, y = eng.kzg_commit_open_batch(key, f, zb)
        assert (C == eng.commit_batch(key, f)).all()
        pf2, y2 = eng.kzg_open_batch(key, f, zb)
        assert (pf == pf2).all() and (y == y2).all()
    epf, ey, ok = orc.kzg_prove(srs, f[0], zb[0])
    assert ok and (pf[0] == epf).all() and (y[0] == ey).all() and (C[0] == orc.msm(srs, f[0])).all()
    key.free()


def test_point_equal_to_key_size_is_fenced(eng):
    """quirk Q2: point == size takes the in-domain branch and indexes out of bounds in the reference"""
    from verkle_kzg_b200 import VkzgError
    N = 16
    srs, key = _key(eng, N, 8)
    f = orc.rand_fr_buf(np.random.default_rng(3), N).reshape(1, N, 32)
    with pytest.raises(VkzgError) as ei:
        eng.kzg_open_batch(key, f, orc.fr_to_buf([N]))
    assert ei.value.status == -3
    _, _, ok = orc.kzg_prove(srs, f[0], orc.fr_to_buf([N])[0])
    assert not ok
    key.free()


def test_trait_surface_kzg(eng):
    from verkle_kzg_b200.vector_commit import KZG, LagrangeBasis, OutOfDomain, fr_from_int
    N = 16
    srs = orc.kzg_setup(N, TAU)
    key = KZG.setup(eng, srs, window_bits=8)
    rng = np.random.default_rng(9)
    data = LagrangeBasis.from_vec_and_domain(orc.rand_fr_buf(rng, 8), N)  # kzg/mod.rs:272
    C = KZG.commit(key, data)
    assert (C == orc.msm(srs, data.evaluations)).all()
    for idx in (0, 3, 7, 17):
        pf = KZG.prove(key, C, idx, data)
        epf, ey, ok = orc.kzg_prove(srs, data.evaluations, orc.fr_to_buf([idx])[0])
        assert ok and (pf["proof"] == epf).all() and (pf["y"] == ey).all()
    with pytest.raises(OutOfDomain):
        KZG.prove(key, C, 16, LagrangeBasis.from_vec(orc.rand_fr_buf(rng, 16)))
    with pytest.raises(NotImplementedError):
        KZG.verify(key, C, 0, pf)
    key.free()


def test_open_batch_property_full_width(eng):
    """size-independent property on a larger batch: q(X) (X - z) == f(X) - y at X = tau, checked on the scalar side:
    [q(tau)] G == proof  and  q(tau) (tau - z) == f(tau) - y, with f(tau) from the commitment side."""
    N = 256
    srs, key = _key(eng, N, 0)
    rng = np.random.default_rng(12)
    B = 200
    f = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    pts = [int(rng.integers(0, N)) for _ in range(B - 20)] + orc.rand_fr(rng, 20)
    zb = orc.fr_to_buf(pts)
    proof, y = eng.kzg_open_batch(key, f, zb)
    C = eng.commit_batch(key, f)
    for i in range(0, B, 9):
        assert orc.kzg_verify_tau(srs, TAU, C[i], zb[i], proof[i], y[i])
    key.free()


def test_fr_vector_ops(eng):
    """L1 / I2: LagrangeBasis AddAssign / Sub / Mul<F>, elementwise_mul, vec_add_and_distribute"""
    rng = np.random.default_rng(21)
    n = 1000
    a, b = orc.rand_fr_buf(rng, n), orc.rand_fr_buf(rng, n)
    a[0], b[0] = orc.fr_to_buf([orc.R_MOD - 1])[0], orc.fr_to_buf([orc.R_MOD - 1])[0]
    x = orc.rand_fr_buf(rng, 1)[0]
    xs = np.tile(x, (n, 1))
    assert (eng.fr_vector_op("add", a, b) == orc.field_op(0, "add", a, b)).all()
    assert (eng.fr_vector_op("sub", a, b) == orc.field_op(0, "sub", a, b)).all()
    assert (eng.fr_vector_op("mul", a, b) == orc.field_op(0, "mul", a, b)).all()
    assert (eng.fr_vector_op("scale", a, x=x) == orc.field_op(0, "mul", a, xs)).all()
    assert (eng.fr_vector_op("axpy", a, b, x) == orc.field_op(0, "add", a, orc.field_op(0, "mul", b, xs))).all()


@pytest.mark.parametrize("m", [1, 5, 16, 20, 256])
def test_kzg_setup_group_ifft(eng, m):
    """KZG::setup (kzg/mod.rs:115-124): powers of tau -> Lagrange SRS, against the oracle's scalar-side construction"""
    gen = orc.g1_generator()
    gkey = eng.load_key(gen.reshape(1, 64), window_bits=8)
    tau = orc.fr_to_buf([TAU])[0]
    powers = eng.kzg_powers(gkey, tau, m)
    exp_pow = orc.g1_mul_gen_batch(orc.fr_to_buf([pow(TAU, i, orc.R_MOD) for i in range(m)]))
    assert (powers == exp_pow).all()
    got = eng.kzg_setup(powers)
    assert (got == orc.kzg_setup(m, TAU)).all()
    # a random secret as well
    t2 = orc.rand_fr(np.random.default_rng(m), 1)[0]
    p2 = orc.g1_mul_gen_batch(orc.fr_to_buf([pow(t2, i, orc.R_MOD) for i in range(m)]))
    assert (eng.kzg_setup(p2) == orc.kzg_setup(m, t2)).all()
    # the closed-form path (the generator's secret is known): the same canonical points
    assert (eng.kzg_setup_from_secret(gkey, tau, m) == got).all()
    assert (eng.kzg_setup_from_secret(gkey, orc.fr_to_buf([t2])[0], m) == orc.kzg_setup(m, t2)).all()
    gkey.free()


def test_kzg_setup_from_secret_edge_cases(eng):
    """tau a domain element (tau w^-j = 1 for one j: the m/n branch), tau = 0 and tau = 1"""
    gen = orc.g1_generator()
    gkey = eng.load_key(gen.reshape(1, 64), window_bits=8)
    w32 = orc.buf_to_fr(orc.domain_gen(32))[0]
    for t, m in ((pow(w32, 5, orc.R_MOD), 20), (pow(w32, 5, orc.R_MOD), 32), (0, 9), (1, 16), (1, 11)):
        assert (eng.kzg_setup_from_secret(gkey, orc.fr_to_buf([t])[0], m) == orc.kzg_setup(m, t)).all(), (t, m)
    gkey.free()


@pytest.mark.parametrize("N,length,domain,wb,B", [(16, 8, 16, 8, 3), (32, 20, 0, 8, 2), (256, 256, 0, 0, 2), (64, 64, 0, 8, 300)])
def test_prove_all_points(eng, N, length, domain, wb, B):
    """KZG::prove_all_points (kzg/mod.rs:200-235) by its contract — the shape of the reference's unregistered test_amortized_proof
    (kzg/mod.rs:298-308): entry i is the single-point proof at i.  Checked against the oracle's prove_point, the tau = 100
    pairing-free verification, and (B = 300: several 1 GiB pieces would need B > 4096 at this width, so one piece) the batched
    single-point entry."""
    srs, key = _key(eng, N, wb)
    rng = np.random.default_rng(N + length + B)
    f = orc.rand_fr_buf(rng, B * length).reshape(B, length, 32)
    want = domain if domain else length
    dn = 1
    while dn < want:
        dn <<= 1
    proof, y = eng.kzg_prove_all_batch(key, f, domain_n=domain)
    assert proof.shape == (B, dn, 64) and y.shape == (B, dn, 32)
    idx = orc.fr_to_buf(list(range(dn)))
    rows = range(B) if B <= 3 else (0, B // 2, B - 1)
    C = eng.commit_batch(key, f)
    for b in rows:
        pf1, y1 = eng.kzg_open_batch(key, np.repeat(f[b:b + 1], dn, axis=0), idx, domain_n=domain)
        assert (proof[b] == pf1).all() and (y[b] == y1).all()
        for i in range(dn):
            assert (y[b, i] == (f[b, i] if i < length else 0)).all()
            eq = orc.divide_by_vanishing(N, f[b], want, i)
            if dn <= 32 or i in (0, 1, dn // 2, dn - 1):
                assert (proof[b, i] == orc.msm(srs, eq)).all(), (b, i)
                if dn == N:
                    assert orc.kzg_verify_tau(srs, TAU, C[b], idx[i], proof[b, i], y[b, i]), (b, i)
    if B > 3:   # every row against the batched single-point path
        pf_all, y_all = eng.kzg_open_batch(key, np.repeat(f, dn, axis=0), np.tile(idx, (B, 1)), domain_n=domain)
        assert (proof.reshape(-1, 64) == pf_all).all() and (y.reshape(-1, 32) == y_all).all()
    key.free()


def test_prove_all_points_errors(eng):
    srs, key = _key(eng, 16, 8)
    f = orc.rand_fr_buf(np.random.default_rng(3), 2 * 20).reshape(2, 20, 32)
    from verkle_kzg_b200._lib import VkzgError
    with pytest.raises(VkzgError):
        eng.kzg_prove_all_batch(key, f)          # rows longer than the key
    key.free()
