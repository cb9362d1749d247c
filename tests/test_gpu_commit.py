"""M1 / D1 parity on the GPU: libvkzg (through the C ABI) against the CPU oracle, compared as canonical
affine bytes (a canonical point has one encoding, so equality of the 64-byte buffers is bit-exactness of
the compressed serialisation too)."""
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


def _bases(n, seed):
    rng = np.random.default_rng(seed)
    k0, k1 = orc.rand_fr(rng, 2)
    return orc.points_walk(k0, k1, n)


@pytest.fixture(scope="module")
def key257(eng):
    bases = _bases(257, 1)
    key = eng.load_key(bases[:256], q=bases[256])
    return bases, key


def test_table_spot_checks(eng):
    """window tables at small width so that every entry can be checked: commit of unit-ish scalars"""
    bases = _bases(4, 7)
    key = eng.load_key(bases, window_bits=8)
    # scalars chosen to hit digit magnitudes 1, 127, 128 (negative), carries, and the top window
    vals = [1, 127, 128, 129, 255, 256, 2 ** 64 - 1, orc.R_MOD - 1, orc.R_MOD - 2, (orc.R_MOD - 1) // 2, 0x8080808080808080, 0]
    for v in vals:
        s = orc.fr_to_buf([v, 0, 0, 0]).reshape(1, 4, 32)
        got = eng.commit_batch(key, s)[0]
        exp = orc.g1_mul(bases[0], orc.fr_to_buf([v])[0])
        assert (got == exp).all(), hex(v)
    key.free()


@pytest.mark.parametrize("w", [1, 4, 31, 32, 129, 256])
def test_commit_batch_matches_oracle(eng, key257, w):
    bases, key = key257
    rng = np.random.default_rng(100 + w)
    B = 5
    s = orc.rand_fr_buf(rng, B * w).reshape(B, w, 32)
    got = eng.commit_batch(key, s)
    exp = orc.commit_batch(bases[:256], s)
    assert (got == exp).all()


def test_commit_edge_scalars(eng, key257):
    bases, key = key257
    edge = [0, 1, 2, orc.R_MOD - 1, orc.R_MOD - 2, 2 ** 128 - 1, 2 ** 128, 2 ** 253, 0x7fff, 0x8000, 0x8001, 0xffff, 0x10000,
            (1 << 240) - 1, int("8000" * 15, 16), int("7fff" * 15, 16)]
    s = orc.fr_to_buf(edge + [0] * (256 - len(edge))).reshape(1, 256, 32)
    zero = np.zeros((1, 256, 32), dtype=np.uint8)
    both = np.concatenate([s, zero])
    got = eng.commit_batch(key, both)
    exp = orc.commit_batch(bases[:256], both)
    assert (got == exp).all()
    assert not got[1].any()  # commitment to the zero vector is the identity


def test_commit_linearity_full_batch(eng, key257):
    """size-independent property at a larger batch: commit(a) + commit(b) == commit(a + b)"""
    bases, key = key257
    rng = np.random.default_rng(5)
    B = 512
    a = orc.rand_fr_buf(rng, B * 256)
    b = orc.rand_fr_buf(rng, B * 256)
    ab = orc.field_op(0, "add", a, b)
    ca = eng.commit_batch(key, a.reshape(B, 256, 32))
    cb = eng.commit_batch(key, b.reshape(B, 256, 32))
    cab = eng.commit_batch(key, ab.reshape(B, 256, 32))
    for i in range(0, B, 37):
        assert (orc.g1_add(ca[i], cb[i]) == cab[i]).all()
    # a few rows against the oracle directly
    exp = orc.commit_batch(bases[:256], a.reshape(B, 256, 32)[:3])
    assert (ca[:3] == exp).all()


def test_commit_width_out_of_range(eng, key257):
    from verkle_kzg_b200 import VkzgError
    _, key = key257
    s = np.zeros((1, 257, 32), dtype=np.uint8)
    with pytest.raises(VkzgError):
        eng.commit_batch(key, s)


def test_to_data_item(eng, key257):
    bases, _ = key257
    pts = np.concatenate([bases[:40], np.zeros((1, 64), dtype=np.uint8)])
    assert (eng.to_data_item(pts) == orc.to_data_item(pts)).all()


def test_g1_sum(eng, key257):
    bases, _ = key257
    acc = np.zeros(64, dtype=np.uint8)
    for p in bases[:19]:
        acc = orc.g1_add(acc, p)
    assert (eng.g1_sum(bases[:19]) == acc).all()
    # P + (-P) + inf
    trio = np.stack([bases[0], orc.g1_neg(bases[0]), np.zeros(64, dtype=np.uint8)])
    assert not eng.g1_sum(trio).any()
    # doubling inside the sum
    assert (eng.g1_sum(np.stack([bases[3], bases[3]])) == orc.g1_add(bases[3], bases[3])).all()


@pytest.mark.parametrize("n,c", [(1, 0), (5, 0), (1000, 0), (1 << 14, 0), (1 << 16, 0), (3000, 9)])
def test_msm_matches_oracle(eng, n, c):
    bases = _bases(n, 11 + n)
    key = eng.load_key(bases, kind=2, window_bits=c)
    rng = np.random.default_rng(n)
    s = orc.rand_fr_buf(rng, n)
    got = eng.msm(key, s)
    exp = orc.msm(bases, s, mode="pippenger")
    assert (got == exp).all()
    if n <= 1000:
        assert (got == orc.msm(bases, s, mode="naive")).all()
    # zip truncation (quirk Q1): fewer scalars than bases
    if n >= 5:
        got2 = eng.msm(key, s[: n // 2])
        assert (got2 == orc.msm(bases[: n // 2], s[: n // 2], mode="pippenger")).all()
    key.free()


@pytest.mark.parametrize("c", [5, 14, 15, 17, 19])
def test_window_counts_at_the_top_of_the_scalar_range(eng, c):
    """W = ceil(255 / c) digits (the scalars are < r < 2^254): widths whose last window is full (15, 17: c W = 255), nearly
    empty (14: two bits) or odd; scalars that maximise the top digit and the carries into it — MSM keys and window keys"""
    n = 600
    bases = _bases(n, 700 + c)
    edge = [orc.R_MOD - 1, orc.R_MOD - 2, (1 << 253), (1 << 253) - 1, (1 << 253) + (1 << 252), orc.R_MOD >> 1, (1 << (c - 1)), (1 << (c - 1)) - 1,
            (1 << c) - 1, sum(1 << (c * w + c - 1) for w in range(255 // c) if c * w + c - 1 < 253), 0, 1]
    rng = np.random.default_rng(c)
    s = np.concatenate([orc.fr_to_buf(edge), orc.rand_fr_buf(rng, n - len(edge))])
    exp = orc.msm(bases, s, mode="pippenger")
    mk = eng.load_key(bases, kind=2, window_bits=c)
    assert (eng.msm(mk, s) == exp).all()
    mk.free()
    if c <= 17:  # window key: 600 bases x W x 2^(c-1) entries
        wk = eng.load_key(bases[:64], window_bits=c)
        got = eng.commit_batch(wk, s[:64][None])
        assert (got[0] == orc.msm(bases[:64], s[:64], mode="pippenger")).all()
        wk.free()


def test_msm_degenerate_scalars(eng):
    n = 4096
    bases = _bases(n, 99)
    key = eng.load_key(bases, kind=2)
    same = np.tile(orc.fr_to_buf([0x1234567])[0], (n, 1))
    assert (eng.msm(key, same) == orc.msm(bases, same, mode="pippenger")).all()
    zero = np.zeros((n, 32), dtype=np.uint8)
    assert not eng.msm(key, zero).any()
    minus1 = np.tile(orc.fr_to_buf([orc.R_MOD - 1])[0], (n, 1))
    assert (eng.msm(key, minus1) == orc.msm(bases, minus1, mode="pippenger")).all()
    key.free()


def test_empty_and_degenerate_shapes(eng, key257):
    """B = 0 and n = 0 are no-ops; a key that is not a power of two still commits (IPA / quotients refuse it)"""
    from verkle_kzg_b200 import VkzgError
    bases, key = key257
    assert eng.commit_batch(key, np.zeros((0, 256, 32), dtype=np.uint8)).shape == (0, 64)
    assert eng.to_data_item(np.zeros((0, 64), dtype=np.uint8)).shape == (0, 32)
    assert not eng.g1_sum(np.zeros((0, 64), dtype=np.uint8)).any()
    rng = np.random.default_rng(3)
    k5 = eng.load_key(bases[:5], q=bases[5], window_bits=8)          # width 5: not a power of two
    s = orc.rand_fr_buf(rng, 2 * 5).reshape(2, 5, 32)
    assert (eng.commit_batch(k5, s) == orc.commit_batch(bases[:5], s)).all()
    with pytest.raises(VkzgError) as ei:                              # the reference's split() breaks on odd lengths
        eng.ipa_prove_batch(k5, s, orc.fr_to_buf([1, 2]), eng.commit_batch(k5, s))
    assert ei.value.status == -4
    k5.free()
    mk = eng.load_key(bases[:8], kind=2, window_bits=8)
    assert not eng.msm(mk, np.zeros((0, 32), dtype=np.uint8)).any()  # empty MSM = identity
    with pytest.raises(VkzgError):
        eng.msm(mk, np.zeros((9, 32), dtype=np.uint8))                # more scalars than bases in the key
    mk.free()


def test_commit_with_identity_bases(eng):
    """a key may contain the point at infinity (e.g. a zero Lagrange commitment): its terms vanish"""
    rng = np.random.default_rng(4)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 8)
    bases[2] = 0
    bases[7] = 0
    key = eng.load_key(bases, window_bits=8)
    s = orc.rand_fr_buf(rng, 3 * 8).reshape(3, 8, 32)
    assert (eng.commit_batch(key, s) == orc.commit_batch(bases, s)).all()
    key.free()
    mk = eng.load_key(bases, kind=2, window_bits=8)
    assert (eng.msm(mk, s[0]) == orc.msm(bases, s[0])).all()
    mk.free()


def test_large_window_bits_and_sparse_scalars(eng):
    """c = 18 tables (more entries per row than c = 16) and sparse / small scalars (many zero digits)"""
    rng = np.random.default_rng(6)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 3)
    key = eng.load_key(bases, window_bits=18)
    vals = [[1, 0, 2 ** 128 - 1], [0, 0, 0], [2 ** 17, 2 ** 18 - 1, 2 ** 35], [orc.R_MOD - 1, 2 ** 253, 12345]]
    s = np.stack([orc.fr_to_buf(v) for v in vals])
    assert (eng.commit_batch(key, s) == orc.commit_batch(bases, s)).all()
    key.free()


def test_group_law_against_public_eip196_vectors(eng):
    """libvkzg's point addition (vkzg_g1_sum) and scalar multiplication (window table of the vector's own point + commit;
    MSM key) on public alt_bn128 vectors that no part of this repository produced (EIP-196, go-ethereum "chfast1")"""
    from test_oracle_primitives import EIP196_ADD, EIP196_MUL, EIP196_2G, eip196_point
    a, b, s = (eip196_point(p) for p in EIP196_ADD)
    assert (eng.g1_sum(np.stack([a, b])) == s).all()
    p, k, r = eip196_point(EIP196_MUL[0]), orc.fr_to_buf([EIP196_MUL[1]]), eip196_point(EIP196_MUL[2])
    for wb in (8, 13, 16):
        key = eng.load_key(p[None], window_bits=wb)
        assert (eng.commit_batch(key, k.reshape(1, 1, 32))[0] == r).all(), wb
        key.free()
    mk = eng.load_key(p[None], kind=2, window_bits=8)
    assert (eng.msm(mk, k) == r).all()
    mk.free()
    g = orc.g1_generator()
    assert (eng.g1_sum(np.stack([g, g])) == eip196_point(EIP196_2G)).all()


def test_msm_dev_only_enqueues_and_can_be_captured_in_a_cuda_graph():
    """vkzg.h: `_dev` calls enqueue on the context's stream and return without synchronising.  vkzg_msm_dev decides between
    the optimistic single pass and the exact counting sort ON THE DEVICE, so the whole call can be captured in a CUDA graph
    and replayed on new scalars — uniform ones (optimistic pass) and degenerate ones (fallback) through the SAME graph."""
    import torch
    from verkle_kzg_b200 import Engine
    n = 1 << 13
    bases = _bases(n, 4242)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng = Engine(0, stream=st.cuda_stream)
        key = eng.load_key(bases, kind=2)
        rng = np.random.default_rng(77)
        s1, s2 = orc.rand_fr_buf(rng, n), orc.rand_fr_buf(rng, n)
        same = np.tile(orc.fr_to_buf([0xabcdef12345])[0], (n, 1))
        d_s = torch.from_numpy(s1).cuda()
        d_out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        eng.msm_dev(key, d_s, n, d_out)            # warm-up: scratch buffers enter the pool
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            eng.msm_dev(key, d_s, n, d_out)
        for s in (s2, same, s1):
            d_s.copy_(torch.from_numpy(s))
            g.replay()
            st.synchronize()
            assert (d_out.cpu().numpy() == orc.msm(bases, s, mode="pippenger")).all()
        key.free()
        eng.close()
