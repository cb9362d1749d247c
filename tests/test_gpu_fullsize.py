"""Size-independent properties at BASELINE.json's full sizes (the oracle cannot finish these in seconds):
   configs[1]  2^14 width-256 IPA proofs: every proof verifies on the device, commit linearity, a sample equals the oracle
   configs[2]  multiproof over 2^12 openings verifies; tampering fails
   configs[3]  2^20-point MSM: linearity  msm(s1) + msm(s2) == msm(s1 + s2),  msm(k * s) == k * msm(s)
   configs[4]  2^17-key tree: native incremental recommit == fresh full commit == level-list commit"""
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


def test_ipa_batch_2_14(eng):
    rng = np.random.default_rng(2014)
    N, B = 256, 1 << 14
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N])
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    zb = orc.fr_to_buf([int(v) for v in rng.integers(0, N, B)])
    C, L, R, tip, y = eng.ipa_commit_prove_batch(key, a, zb)
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
    # y is the opened evaluation a[z]
    zi = np.array([int.from_bytes(bytes(r), "little") for r in orc.from_mont(0, zb)])
    assert (y == a[np.arange(B), zi]).all()
    # a sample against the oracle, byte for byte
    for i in (0, 8191, 8192, B - 1):      # both half-batches of the two-stream split
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all()
    # tampering one proof flips only that verdict
    tip2 = tip.copy()
    tip2[12345] = tip[0]
    ok = eng.ipa_verify_batch(key, zb, C, L, R, tip2, y)
    assert not ok[12345] and ok.sum() == B - 1
    key.free()


def test_multiproof_2_12(eng):
    rng = np.random.default_rng(2012)
    N, m = 256, 1 << 12
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N])
    f = orc.rand_fr_buf(rng, m * N).reshape(m, N, 32)
    C = eng.commit_batch(key, f)
    z = rng.integers(0, N, m).astype(np.uint64)
    yq = f[np.arange(m), z.astype(np.int64)]
    mp = eng.multiproof_prove(key, "ipa", f, C, z, yq)
    assert eng.multiproof_verify_ipa(key, C, z, yq, mp)
    y2 = yq.copy()
    y2[4000] = yq[0]
    assert not eng.multiproof_verify_ipa(key, C, z, y2, mp)
    key.free()


def test_msm_2_20_linearity(eng):
    import torch
    n = 1 << 20
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20)
    import bench
    pts = bench.make_points_dev(torch, eng, n, gen).cpu().numpy()
    key = eng.load_key(pts, kind=2)
    rng = np.random.default_rng(20)
    s1, s2 = orc.rand_fr_buf(rng, n), orc.rand_fr_buf(rng, n)
    m1, m2 = eng.msm(key, s1), eng.msm(key, s2)
    assert (eng.msm(key, orc.field_op(0, "add", s1, s2)) == orc.g1_add(m1, m2)).all()
    kk = orc.rand_fr_buf(rng, 1)
    assert (eng.msm(key, orc.field_op(0, "mul", s1, np.tile(kk, (n, 1)))) == orc.g1_mul(m1, kk[0])).all()
    # a prefix against the oracle's bucket method
    assert (eng.msm(key, s1[:5000]) == orc.msm(pts[:5000], s1[:5000], mode="pippenger")).all()
    # the whole MSM (n >= 2^19: two bucket passes, the second scatter overlapped with the first pass) against the sum of
    # range MSMs small enough for the single-pass form, cut at points that do not line up with the split
    d_s = torch.from_numpy(s1).cuda()
    d_out = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
    acc = np.zeros(64, dtype=np.uint8)
    cuts = [0, 200_000, 262_144 + 77, 600_001, 850_000, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        eng.msm_dev(key, d_s[a:b], b - a, d_out, first=a)
        eng.sync()
        acc = orc.g1_add(acc, d_out.cpu().numpy()[0])
    assert (acc == m1).all()
    # one scalar changed in each part of the split: the result moves by exactly (s' - s) P_j
    for j in (5, n // 4 - 1, n // 4 + 3, n - 1):
        t = s1.copy()
        t[j] = s2[j]
        diff = orc.field_op(0, "sub", s2[j:j + 1], s1[j:j + 1])[0]
        assert (eng.msm(key, t) == orc.g1_add(m1, orc.g1_mul(pts[j], diff))).all(), j
    key.free()


def test_tree_2_17_incremental_equals_full(eng):
    from verkle_kzg_b200.tree import NativeVerkleTree, build_levels
    rng = np.random.default_rng(2017)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 256)
    key = eng.load_key(bases)
    n = 1 << 17
    keys = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    vals = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    t = NativeVerkleTree(32, 256)
    half = n // 2
    t.insert_many(keys[:half], vals[:half])
    t.commitment(eng, key)
    t.insert_many(keys[half:], vals[half:])           # second bulk insert: only dirty paths are recommitted
    inc = t.commitment(eng, key)
    fresh = NativeVerkleTree(32, 256)
    fresh.insert_many(keys, vals)
    assert (fresh.commitment(eng, key) == inc).all()
    assert t.last_committed < fresh.last_committed
    # the Python mirror's level lists through vkzg_tree_commit_levels give the same root
    sub = 4096
    small = NativeVerkleTree(32, 256)
    small.insert_many(keys[:sub], vals[:sub])
    assert (small.commitment(eng, key) == eng.tree_commit_levels(key, build_levels(keys[:sub], vals[:sub], 256))).all()
    assert (NativeVerkleTreeRoot(eng, key, keys[:300], vals[:300]) == orc.tree_commit(bases, keys[:300], vals[:300])).all()
    key.free()


def NativeVerkleTreeRoot(eng, key, keys, vals):
    from verkle_kzg_b200.tree import NativeVerkleTree
    t = NativeVerkleTree(keys.shape[1], 256)
    t.insert_many(keys, vals)
    return t.commitment(eng, key)
