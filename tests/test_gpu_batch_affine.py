"""The batch-affine summation tree (csrc/batch_affine.cu) against the XYZZ path and the oracle: both produce canonical
affine points, so every byte must agree whatever the order of summation.  Exceptional pairs are forced on purpose:
zero digits and zero scalars (identity operands, padding), duplicate bases with equal scalars (tangent case), a base and
its negative with equal scalars (cancellation to the identity), odd list lengths at every level, widths 1 .. 257."""
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


def _both(eng, fn):
    eng.set_option(eng.OPT_BATCH_AFFINE, 0)
    ref = fn()
    eng.set_option(eng.OPT_BATCH_AFFINE, 1)
    try:
        got = fn()
    finally:
        eng.set_option(eng.OPT_BATCH_AFFINE, -1)
    return ref, got


@pytest.mark.parametrize("w,wb", [(1, 8), (2, 8), (3, 16), (37, 12), (128, 16), (256, 16), (256, 13)])
def test_commits_match_xyzz_path_and_oracle(eng, w, wb):
    rng = np.random.default_rng(1000 + w + wb)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, max(w, 2))
    key = eng.load_key(bases, window_bits=wb)
    B = 70
    s = orc.rand_fr_buf(rng, B * w).reshape(B, w, 32)
    s[3] = 0                                                   # zero vector: every entry is padding
    s[4, : w // 2] = 0                                         # half the terms vanish
    s[5] = orc.fr_to_buf([orc.R_MOD - 1] * w)                  # digits at their extremes
    s[6] = orc.fr_to_buf([1] * w)                              # a single non-zero digit per scalar
    ref, got = _both(eng, lambda: eng.commit_batch(key, s))
    assert (ref == got).all()
    assert (got[:12] == orc.commit_batch(bases[:w], s[:12])).all()
    key.free()


def test_tangent_and_cancellation_pairs(eng):
    """duplicate bases with equal scalars put EQUAL points next to each other (tangent slope), a base next to its negative
    with equal scalars makes opposite points (identity) — at level 0 and, through partial sums, at higher levels"""
    rng = np.random.default_rng(7)
    k0, k1 = orc.rand_fr(rng, 2)
    p = orc.points_walk(k0, k1, 4)
    w = 16
    bases = np.stack([p[0], p[0], p[1], orc.g1_neg(p[1]), p[2], p[2], p[2], p[2],
                      p[3], orc.g1_neg(p[3]), p[3], orc.g1_neg(p[3]), p[0], p[1], p[0], p[1]])
    key = eng.load_key(bases, window_bits=8)
    a, b, c, d = orc.rand_fr(rng, 4)
    rows = [[a, a, b, b, c, c, c, c, d, d, d, d, a, b, a, b],      # pairs equal / opposite at level 0, quads at level 1
            [a] * 16, [0] * 16, [1] * 16, [orc.R_MOD - 1] * 16,
            [a, a, a, a, 0, 0, 0, 0, b, b, b, b, 0, 0, 0, 0]]
    s = np.stack([orc.fr_to_buf(r) for r in rows])
    ref, got = _both(eng, lambda: eng.commit_batch(key, s))
    assert (ref == got).all()
    assert (got == orc.commit_batch(bases, s)).all()
    key.free()


def test_ipa_proofs_match_through_the_tree(eng):
    """the L / R cross terms of every round through the tree: base selection (ipa_m, the Q row) and odd term counts (129)"""
    rng = np.random.default_rng(11)
    N = 256
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N])
    B = 12
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    a[1] = 0
    zb = orc.fr_to_buf([0, 5, N - 1, N + 3] + [int(v) for v in rng.integers(0, N, B - 4)])
    ref, got = _both(eng, lambda: eng.ipa_commit_prove_batch(key, a, zb))
    for r, g in zip(ref, got):
        assert (r == g).all()
    C, L, R, tip, y = got
    for i in (0, 1, 3, B - 1):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all(), i
    key.free()


def test_big_batch_takes_the_tree_by_default_only_when_enabled(eng):
    """size-independent property at a batch that crosses the automatic threshold: commit linearity and equality of the paths"""
    rng = np.random.default_rng(13)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 256)
    key = eng.load_key(bases)
    B = 2048
    s = orc.rand_fr_buf(rng, B * 256).reshape(B, 256, 32)
    ref, got = _both(eng, lambda: eng.commit_batch(key, s))
    assert (ref == got).all()
    assert (got[::500] == orc.commit_batch(bases, s[::500])).all()
    key.free()
