"""I1 / I3 / I4 / B1 parity on the GPU: proofs from libvkzg equal the oracle's restatement of
low_level_ipa byte for byte, verify under the oracle's verifier, and tampered proofs fail under the
device verifier (the reference's own test properties, ipa/mod.rs:382-421)."""
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


def _setup(eng, N, seed, window_bits=0):
    rng = np.random.default_rng(seed)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N], window_bits=window_bits)
    return rng, bases, key


def _points(rng, N, B):
    """in-domain indices, the boundary values, and out-of-domain points"""
    vals = [int(rng.integers(0, N)) for _ in range(B)]
    special = [0, N - 1, N, N + 1, 2 * N, orc.R_MOD - 2] + orc.rand_fr(rng, 2)
    for i, v in enumerate(special[:B]):
        vals[i] = v
    return vals


@pytest.mark.parametrize("N,wb", [(256, 0), (32, 12), (2, 8), (4, 8)])
def test_prove_matches_oracle_and_verifies(eng, N, wb):
    rng, bases, key = _setup(eng, N, 40 + N, wb)
    B = 10
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    C = eng.commit_batch(key, a)
    zs = _points(rng, N, B)
    zb = orc.fr_to_buf(zs)
    L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C)
    for i in range(B):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all(), f"proof {i} (z={zs[i]})"
        assert (tip[i] == etip).all() and (y[i] == ey).all()
        assert orc.ipa_verify(bases, N, C[i], zb[i], L[i], R[i], tip[i], y[i])
    ok = eng.ipa_verify_batch(key, zb, C, L, R, tip, y)
    assert ok.all()
    # tamper: y, tip, one L, the point — each must fail on the device verifier and on the oracle's
    one = orc.fr_to_buf([1])[0]
    y2 = y.copy()
    y2[0] = orc.field_op(0, "add", y[0], one)[0]
    tip2 = tip.copy()
    tip2[1] = orc.field_op(0, "add", tip[1], one)[0]
    L2 = L.copy()
    L2[2, 0] = bases[1]
    z2 = zb.copy()
    z2[3] = orc.field_op(0, "add", zb[3], one)[0]
    assert list(eng.ipa_verify_batch(key, zb, C, L, R, tip, y2)) == [False] + [True] * (B - 1)
    assert list(eng.ipa_verify_batch(key, zb, C, L, R, tip2, y)) == [True, False] + [True] * (B - 2)
    assert list(eng.ipa_verify_batch(key, zb, C, L2, R, tip, y)) == [True, True, False] + [True] * (B - 3)
    assert list(eng.ipa_verify_batch(key, z2, C, L, R, tip, y)) == [True, True, True, False] + [True] * (B - 4)
    assert not orc.ipa_verify(bases, N, C[0], zb[0], L[0], R[0], tip[0], y2[0])
    key.free()


def test_prove_with_inflight_transcript(eng):
    """prove_point continuing a caller's transcript (lib.rs:127-133 / multiproof.rs:174)"""
    N = 32
    rng, bases, key = _setup(eng, N, 77, 12)
    a = orc.rand_fr_buf(rng, N).reshape(1, N, 32)
    C = eng.commit_batch(key, a)
    zb = orc.fr_to_buf([orc.rand_fr(rng, 1)[0]])
    prefix = bytes(rng.integers(0, 256, 66, dtype=np.uint8))
    L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C, prefix=prefix, dst="multiproof")
    eL, eR, etip, ey = orc.ipa_prove(bases, N, a[0], C[0], zb[0], prefix=prefix, dst="multiproof")
    assert (L[0] == eL).all() and (R[0] == eR).all() and (tip[0] == etip).all() and (y[0] == ey).all()
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y, prefix=prefix, dst="multiproof").all()
    assert not eng.ipa_verify_batch(key, zb, C, L, R, tip, y).any()  # different transcript -> different challenges
    key.free()


def test_special_vectors(eng):
    """zero vector (identity commitment, L = R = identity), unit vector, r+i ramp (benches/ipa.rs:54-62)"""
    N = 32
    rng, bases, key = _setup(eng, N, 78, 12)
    r0 = orc.rand_fr(rng, 1)[0]
    vecs = [[0] * N, [1] + [0] * (N - 1), [(r0 + i) % orc.R_MOD for i in range(N)], [orc.R_MOD - 1] * N]
    a = np.stack([orc.fr_to_buf(v) for v in vecs])
    C = eng.commit_batch(key, a)
    zb = orc.fr_to_buf([3, 0, N + 5, 17])
    L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C)
    for i in range(len(vecs)):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all(), i
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
    key.free()


def test_prove_commitment(eng):
    N = 32
    rng, bases, key = _setup(eng, N, 79, 12)
    B = 3
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    C = eng.commit_batch(key, a)
    L, R, tip = eng.ipa_prove_commitment_batch(key, a, C)
    for i in range(B):
        eL, eR, etip = orc.ipa_prove_commitment(bases, N, a[i], C[i])
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all()
        assert orc.ipa_verify_commitment(bases, N, C[i], L[i], R[i], tip[i])
    key.free()


def test_verify_commitment_proof_on_the_device(eng):
    """IPA::verify_commitment_proof (ipa/mod.rs:238-265) on the device: accepts what prove_commitment made (on a key WITHOUT
    Q too), rejects a tampered tip / L / commitment — the same verdicts as the oracle's verifier"""
    N = 32
    rng, bases, key = _setup(eng, N, 83, 12)
    B = 5
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    C = eng.commit_batch(key, a)
    L, R, tip = eng.ipa_prove_commitment_batch(key, a, C)
    assert eng.ipa_verify_commitment_batch(key, C, L, R, tip).all()
    one = orc.fr_to_buf([1])[0]
    tip2, L2, C2 = tip.copy(), L.copy(), C.copy()
    tip2[0] = orc.field_op(0, "add", tip[0], one)[0]
    L2[1, 2] = bases[3]
    C2[2] = bases[4]
    assert list(eng.ipa_verify_commitment_batch(key, C, L, R, tip2)) == [False, True, True, True, True]
    assert list(eng.ipa_verify_commitment_batch(key, C, L2, R, tip)) == [True, False, True, True, True]
    assert list(eng.ipa_verify_commitment_batch(key, C2, L, R, tip)) == [True, True, False, True, True]
    assert not orc.ipa_verify_commitment(bases, N, C[0], L[0], R[0], tip2[0])
    assert not orc.ipa_verify_commitment(bases, N, C2[2], L[2], R[2], tip[2])
    key.free()
    noq = eng.load_key(bases[:N], window_bits=12)
    assert eng.ipa_verify_commitment_batch(noq, C, L, R, tip).all()
    noq.free()
    # through the trait mirror
    from verkle_kzg_b200.vector_commit import IPA, LagrangeBasis
    up = IPA.setup(eng, bases[:N], q=bases[N], window_bits=12)
    data = LagrangeBasis.from_vec(a[0])
    c0 = IPA.commit(up, data)
    pr = IPA.prove_commitment(up, c0, data)
    assert IPA.verify_commitment_proof(up, c0, pr)
    pr["tip"] = tip2[0]
    assert not IPA.verify_commitment_proof(up, c0, pr)
    up.free()


@pytest.mark.parametrize("plen", [0, 15, 16, 17, 160, 161, 208, 209, 1000, 4099])
def test_inflight_transcript_of_any_length(eng, plen):
    """the reference's transcript has no size limit (transcript.rs:34-52): prefixes beyond the 160 inline bytes are
    pre-hashed block-wise on the host (SHA-256 midstate) — every block-boundary case against the oracle"""
    N = 4
    rng, bases, key = _setup(eng, N, 84 + plen, 8)
    a = orc.rand_fr_buf(rng, 2 * N).reshape(2, N, 32)
    C = eng.commit_batch(key, a)
    zb = orc.fr_to_buf([1, N + 3])
    prefix = bytes(rng.integers(0, 256, plen, dtype=np.uint8))
    L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C, prefix=prefix, dst="multiproof")
    for i in range(2):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i], prefix=prefix, dst="multiproof")
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all(), (plen, i)
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y, prefix=prefix, dst="multiproof").all()
    if plen:
        other = bytes([prefix[0] ^ 1]) + prefix[1:]
        assert not eng.ipa_verify_batch(key, zb, C, L, R, tip, y, prefix=other, dst="multiproof").any()
    key.free()


def test_barycentric(eng):
    N = 256
    rng, bases, key = _setup(eng, N, 80)
    zs = [0, 5, N - 1, N, N + 1, orc.R_MOD - 2] + orc.rand_fr(rng, 3)
    zb = orc.fr_to_buf(zs)
    got = eng.barycentric_batch(key, zb)
    for i in range(len(zs)):
        assert (got[i] == orc.barycentric(N, zb[i])).all(), zs[i]
    key.free()


def test_batch_roundtrip_full_width(eng):
    """size-independent property at a bigger batch: every proof of a 256-wide batch verifies on the device,
    a sample equals the oracle"""
    N = 256
    rng, bases, key = _setup(eng, N, 81)
    B = 256
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    C = eng.commit_batch(key, a)
    zb = orc.fr_to_buf([int(rng.integers(0, N)) for _ in range(B)])
    L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C)
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
    for i in (0, 100, 255):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all()
    key.free()


def test_commit_prove_in_one_call(eng):
    N = 32
    rng, bases, key = _setup(eng, N, 82, 12)
    B = 6
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    zb = orc.fr_to_buf(_points(rng, N, B))
    C, L, R, tip, y = eng.ipa_commit_prove_batch(key, a, zb)
    C2 = eng.commit_batch(key, a)
    L2, R2, tip2, y2 = eng.ipa_prove_batch(key, a, zb, C2)
    assert (C == C2).all() and (L == L2).all() and (R == R2).all() and (tip == tip2).all() and (y == y2).all()
    assert (C == orc.commit_batch(bases[:N], a)).all()
    key.free()
