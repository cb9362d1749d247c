"""Oracle pinning, part 2: the scheme-level restatement against the independent Python model and the
reference's own test properties."""
import numpy as np
import pytest

import orc
import pyref


def _crs(seed, n):
    """n random-looking affine points: (k0 + i k1) G."""
    rng = np.random.default_rng(seed)
    k0, k1 = orc.rand_fr(rng, 2)
    buf = orc.points_walk(k0, k1, n)
    return buf, orc.buf_to_pts(buf)


@pytest.mark.parametrize("n", [1, 2, 5, 16])
def test_commit_naive_vs_model_and_pippenger(n):
    rng = np.random.default_rng(10 + n)
    bbuf, bpts = _crs(100 + n, n)
    s = orc.rand_fr(rng, n)
    sbuf = orc.fr_to_buf(s)
    want = pyref.inner_product_g(bpts, s)
    assert orc.buf_to_pts(orc.msm(bbuf, sbuf, "naive", 1))[0] == want
    assert orc.buf_to_pts(orc.msm(bbuf, sbuf, "naive", 3))[0] == want
    assert orc.buf_to_pts(orc.msm(bbuf, sbuf, "pippenger", 2))[0] == want


def test_msm_modes_agree_medium_and_edge_scalars():
    rng = np.random.default_rng(11)
    n = 300
    bbuf, _ = _crs(7, n)
    s = orc.rand_fr(rng, n)
    s[0], s[1], s[2], s[3] = 0, 1, orc.R_MOD - 1, 2 ** 253
    sbuf = orc.fr_to_buf(s)
    assert (orc.msm(bbuf, sbuf, "naive", 8) == orc.msm(bbuf, sbuf, "pippenger", 8)).all()
    # zip truncation (quirk Q1): fewer scalars than bases
    assert (orc.msm(bbuf, sbuf[:17], "naive", 1) == orc.msm(bbuf[:17], sbuf[:17], "pippenger", 1)).all()


def test_commit_batch_matches_single():
    rng = np.random.default_rng(12)
    bbuf, _ = _crs(8, 8)
    sc = orc.rand_fr_buf(rng, 3 * 8).reshape(3, 8, 32)
    out = orc.commit_batch(bbuf, sc, 2)
    for k in range(3):
        assert (out[k] == orc.msm(bbuf, sc[k], "naive", 1)).all()


@pytest.mark.parametrize("N", [4, 32])
def test_barycentric_evaluate_quotients_vs_model(N):
    rng = np.random.default_rng(20 + N)
    data = orc.rand_fr(rng, N)
    dbuf = orc.fr_to_buf(data)
    pts = [0, 1, N - 1, N, N + 1, 2 * N, orc.rand_fr(rng, 1)[0]]
    for z in pts:
        zb = orc.fr_to_buf([z])
        assert orc.buf_to_fr(orc.barycentric(N, zb)) == pyref.barycentric(N, z)
        assert orc.buf_to_fr(orc.evaluate(N, dbuf, N, zb))[0] == pyref.evaluate(N, data, N, z)
    # barycentric really evaluates the interpolant: compare against Lagrange interpolation done naively
    z = pts[-1]
    w = pyref.group_gen(N)
    xs = [pow(w, i, pyref.R_MOD) for i in range(N)]
    val = 0
    for i in range(N):
        num = den = 1
        for j in range(N):
            if i != j:
                num = num * (z - xs[j]) % pyref.R_MOD
                den = den * (xs[i] - xs[j]) % pyref.R_MOD
        val = (val + data[i] * num * pow(den, -1, pyref.R_MOD)) % pyref.R_MOD
    assert pyref.evaluate(N, data, N, z) == val
    for idx in [0, 1, N - 1]:
        assert orc.buf_to_fr(orc.divide_by_vanishing(N, dbuf, N, idx)) == pyref.divide_by_vanishing(N, data, N, idx)
    for z in [N + 1, pts[-1]]:
        got = orc.buf_to_fr(orc.divide_by_vanishing_outside(N, dbuf, N, orc.fr_to_buf([z])))
        assert got == pyref.divide_by_vanishing_outside(N, data, N, z)
    # shorter data than the domain (kzg test_single_proof shape: 8 of 16)
    short = data[: N // 2]
    sb = orc.fr_to_buf(short)
    for idx in [0, N // 2, N - 1]:
        assert orc.buf_to_fr(orc.divide_by_vanishing(N, sb, N, idx)) == pyref.divide_by_vanishing(N, short, N, idx)
    ev, inv = orc.vanishing(N)
    pev, pinv = pyref.vanishing_evaluations(N)
    assert orc.buf_to_fr(ev) == pev and orc.buf_to_fr(inv) == pinv


@pytest.mark.parametrize("N", [2, 8])
def test_ipa_prove_matches_model_and_verifies(N):
    rng = np.random.default_rng(30 + N)
    bbuf, bpts = _crs(200 + N, N + 1)
    a = orc.rand_fr(rng, N)
    abuf = orc.fr_to_buf(a)
    C = pyref.inner_product_g(bpts[:N], a)
    Cb = orc.pts_to_buf([C])[0]
    assert (orc.msm(bbuf[:N], abuf, "naive", 1) == Cb).all()
    for z in [1, 2 * N, orc.rand_fr(rng, 1)[0]]:
        zb = orc.fr_to_buf([z])
        L, R, tip, y = orc.ipa_prove(bbuf, N, abuf, Cb, zb)
        mL, mR, mtip, my = pyref.ipa_prove_point(bpts[:N], bpts[N], N, a, C, z)
        assert orc.buf_to_pts(L) == mL and orc.buf_to_pts(R) == mR
        assert orc.buf_to_fr(tip)[0] == mtip and orc.buf_to_fr(y)[0] == my
        assert my == pyref.evaluate(N, a, N, z)
        assert orc.ipa_verify(bbuf, N, Cb, zb, L, R, tip, y)
        assert pyref.ipa_verify_point(bpts[:N], bpts[N], N, C, z, (mL, mR, mtip, my))
        # tamper (ipa/mod.rs:404-421)
        assert not orc.ipa_verify(bbuf, N, orc.g1_add(Cb, orc.g1_generator()), zb, L, R, tip, y)
        bad_y = orc.fr_to_buf([(my + 1) % orc.R_MOD])[0]
        assert not orc.ipa_verify(bbuf, N, Cb, zb, L, R, tip, bad_y)
    # in-flight transcript continuation
    prefix = bytes(range(66))
    L, R, tip, y = orc.ipa_prove(bbuf, N, abuf, Cb, orc.fr_to_buf([3]), prefix=prefix, dst="multiproof")
    mL, mR, mtip, my = pyref.ipa_prove_point(bpts[:N], bpts[N], N, a, C, 3, pyref.Transcript("multiproof", prefix))
    assert orc.buf_to_pts(L) == mL and orc.buf_to_fr(tip)[0] == mtip
    assert orc.ipa_verify(bbuf, N, Cb, orc.fr_to_buf([3]), L, R, tip, y, prefix=prefix, dst="multiproof")
    assert not orc.ipa_verify(bbuf, N, Cb, orc.fr_to_buf([3]), L, R, tip, y)


def test_ipa_reference_test_shapes_size32():
    """ipa/mod.rs:382-421 test_commit_evaluations + test_eval_proof at SIZE = 32 (data = 0..31)."""
    N = 32
    bbuf, _ = _crs(77, N + 1)
    abuf = orc.fr_to_buf(list(range(N)))
    C = orc.msm(bbuf[:N], abuf, "naive", 4)
    L, R, tip = orc.ipa_prove_commitment(bbuf, N, abuf, C)
    assert orc.ipa_verify_commitment(bbuf, N, C, L, R, tip)
    assert not orc.ipa_verify_commitment(bbuf, N, orc.g1_add(C, orc.g1_generator()), L, R, tip)
    idx = 13
    zb = orc.fr_to_buf([idx])
    L, R, tip, y = orc.ipa_prove(bbuf, N, abuf, C, zb)
    assert orc.buf_to_fr(y)[0] == idx
    assert orc.ipa_verify(bbuf, N, C, zb, L, R, tip, y)
    zo = orc.fr_to_buf([2 * N])
    Lo, Ro, tipo, yo = orc.ipa_prove(bbuf, N, abuf, C, zo)
    assert orc.ipa_verify(bbuf, N, C, zo, Lo, Ro, tipo, yo)
    assert not orc.ipa_verify(bbuf, N, C, zb, Lo, Ro, tipo, yo)
    # batch entry point == single
    a2 = np.stack([abuf, abuf])
    Lb, Rb, tb, yb = orc.ipa_prove_batch(bbuf, N, a2, np.stack([C, C]), np.stack([zb[0], zo[0]]), 2)
    assert (Lb[0] == L).all() and (Rb[1] == Ro).all() and (tb[1] == tipo).all() and (yb[0] == y).all()


def test_kzg_setup_prove_verify():
    """kzg/mod.rs:278-297 test_single_proof (DATA_SIZE 8, MAX_CRS 16, tau = 100)."""
    tau, n, dlen = 100, 16, 8
    lag = orc.kzg_setup(n, tau)
    lag_pts = orc.buf_to_pts(lag)
    assert lag_pts[:4] == pyref.kzg_setup(n, tau)[:4]
    # sum of Lagrange commitments = [1]G ; sum w^i L_i = [tau] G
    w = pyref.group_gen(n)
    assert pyref.inner_product_g(lag_pts, [1] * n) == pyref.G1_GEN
    assert pyref.inner_product_g(lag_pts, [pow(w, i, pyref.R_MOD) for i in range(n)]) == pyref.g_mul(pyref.G1_GEN, tau)
    rng = np.random.default_rng(40)
    data = orc.rand_fr(rng, dlen)
    dbuf = orc.fr_to_buf(data)
    C = orc.msm(lag, dbuf, "naive", 1)
    for i in list(range(n)) + [n + 1, orc.rand_fr(rng, 1)[0]]:
        zb = orc.fr_to_buf([i])
        proof, y, ok = orc.kzg_prove(lag, dbuf, zb)
        assert ok
        assert orc.kzg_verify_tau(lag, tau, C, zb, proof, y)
        if dlen <= i < n:
            assert orc.buf_to_fr(y)[0] == 0
        if i in (0, 3, 9, n + 1):
            mp, my = pyref.kzg_prove_point(lag_pts, data, i)
            assert orc.buf_to_pts(proof)[0] == mp and orc.buf_to_fr(y)[0] == my
        bad = orc.fr_to_buf([(orc.buf_to_fr(y)[0] + 1) % orc.R_MOD])[0]
        assert not orc.kzg_verify_tau(lag, tau, C, zb, proof, bad)
    # quirk Q2: point == size indexes out of bounds in the reference -> fenced
    _, _, ok = orc.kzg_prove(lag, dbuf, orc.fr_to_buf([n]))
    assert not ok
    # non power-of-two max_items (benches/kzg.rs: 20 values in a 32-point SRS) pads the domain
    assert len(orc.kzg_setup(20, tau)) == 32


@pytest.mark.parametrize("scheme", ["ipa", "kzg"])
def test_multiproof_matches_model_and_verifies(scheme):
    """multiproof.rs:261-357 at a size the Python model finishes quickly (N = 4, 5 queries)."""
    N, m, tau = 4, 5, 100
    rng = np.random.default_rng(50)
    if scheme == "ipa":
        bbuf, bpts = _crs(300, N + 1)
        cb = bbuf[:N]
    else:
        bbuf = orc.kzg_setup(N, tau)
        bpts = orc.buf_to_pts(bbuf)
        cb = bbuf
    f = [[(r + i) % orc.R_MOD for i in range(N)] for r in orc.rand_fr(rng, m)]
    fbuf = np.stack([orc.fr_to_buf(row) for row in f])
    z = np.array([1, 3, 1, 0, 3], dtype=np.uint64)
    C = orc.commit_batch(cb, fbuf, 2)
    Cp = orc.buf_to_pts(C)
    y = [f[k][int(z[k])] for k in range(m)]
    ybuf = orc.fr_to_buf(y)
    proof = orc.multiproof_prove(scheme, bbuf, N, fbuf, C, z, ybuf)
    mproof, md = pyref.multiproof_prove(scheme, bpts, N, [(f[k], Cp[k], int(z[k]), y[k]) for k in range(m)])
    assert orc.buf_to_pts(proof["D"])[0] == md
    if scheme == "ipa":
        assert orc.buf_to_pts(proof["L"]) == mproof[0] and orc.buf_to_pts(proof["R"]) == mproof[1]
        assert orc.buf_to_fr(proof["tip"])[0] == mproof[2] and orc.buf_to_fr(proof["y"])[0] == mproof[3]
    else:
        assert orc.buf_to_pts(proof["L"])[0] == mproof[0] and orc.buf_to_fr(proof["y"])[0] == mproof[1]
    assert orc.multiproof_verify(scheme, bbuf, N, C, z, ybuf, proof, tau)
    assert pyref.multiproof_verify(scheme, bpts, N, [(Cp[k], int(z[k]), y[k]) for k in range(m)], mproof, md, tau)
    bad = dict(proof)
    bad["D"] = orc.g1_add(proof["D"], orc.g1_generator())
    assert not orc.multiproof_verify(scheme, bbuf, N, C, z, ybuf, bad, tau)
    ybad = ybuf.copy()
    ybad[0] = orc.fr_to_buf([(y[0] + 1) % orc.R_MOD])[0]
    if scheme == "ipa":
        # quirk Q4: verify_multiproof never uses y (g2_of_t is dead), so a wrong y still verifies in the
        # reference for the IPA scheme only through the transcript -> it does change r, hence fails.
        assert not orc.multiproof_verify(scheme, bbuf, N, C, z, ybad, proof, tau)


def test_multiproof_reference_shape_ipa_size32():
    """multiproof.rs:261-308: 20 vectors of width 32 (r + i), random z; oracle prove -> oracle verify."""
    N, m = 32, 20
    rng = np.random.default_rng(51)
    bbuf, _ = _crs(301, N + 1)
    f = [[(r + i) % orc.R_MOD for i in range(N)] for r in orc.rand_fr(rng, m)]
    fbuf = np.stack([orc.fr_to_buf(row) for row in f])
    z = rng.integers(0, N, size=m).astype(np.uint64)
    C = orc.commit_batch(bbuf[:N], fbuf, 8)
    ybuf = orc.fr_to_buf([f[k][int(z[k])] for k in range(m)])
    proof = orc.multiproof_prove("ipa", bbuf, N, fbuf, C, z, ybuf)
    assert orc.multiproof_verify("ipa", bbuf, N, C, z, ybuf, proof)
    bad = dict(proof)
    bad["D"] = orc.g1_add(proof["D"], orc.g1_generator())
    assert not orc.multiproof_verify("ipa", bbuf, N, C, z, ybuf, bad)


def test_tree_commitment_small():
    """node.rs:212-277 on a tiny tree, recomputed by hand with the Python model."""
    rng = np.random.default_rng(60)
    bbuf, bpts = _crs(400, 256)
    keys = np.array([[1, 2, 3], [1, 7, 9], [200, 0, 130]], dtype=np.uint8)
    vals = np.frombuffer(rng.bytes(3 * 32), dtype=np.uint8).reshape(3, 32)
    root = orc.buf_to_pts(orc.tree_commit(bbuf, keys, vals, ext_width=256))[0]

    def ext(key, val):
        idx = int(key[-1])
        lo = int.from_bytes(bytes(val[:16]), "little")
        hi = int.from_bytes(bytes(val[16:]), "little")
        c1, c2 = [0] * 256, [0] * 256
        tgt = c1 if idx < 128 else c2
        tgt[(2 * idx) % 256], tgt[(2 * idx + 1) % 256] = lo, hi
        C1, C2 = pyref.inner_product_g(bpts, c1), pyref.inner_product_g(bpts, c2)
        stem = int.from_bytes(bytes(key), "little") % pyref.R_MOD
        return pyref.inner_product_g(bpts, [1, stem, pyref.to_data_item(C1), pyref.to_data_item(C2)])

    e0, e1, e2 = ext(keys[0], vals[0]), ext(keys[1], vals[1]), ext(keys[2], vals[2])
    inner = [0] * 256
    inner[2], inner[7] = pyref.to_data_item(e0), pyref.to_data_item(e1)
    top = [0] * 256
    top[1] = pyref.to_data_item(pyref.inner_product_g(bpts, inner))
    top[200] = pyref.to_data_item(e2)
    assert root == pyref.inner_product_g(bpts, top)
    # insertion order does not matter; overwrite keeps the last value (lib.rs:305-317)
    root2 = orc.tree_commit(bbuf, keys[::-1].copy(), vals[::-1].copy(), 256)
    assert orc.buf_to_pts(root2)[0] == root
