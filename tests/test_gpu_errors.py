"""Error behaviour of the C ABI on the GPU: status codes instead of crashes for bad arguments, wrong key kinds and
shapes the reference itself rejects or panics on."""
import ctypes

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


def test_status_codes(eng):
    from verkle_kzg_b200 import VkzgError, _lib
    L = _lib.lib()
    rng = np.random.default_rng(1)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 9)
    wkey = eng.load_key(bases[:8], q=bases[8], window_bits=8)
    kkey = eng.load_key(bases[:8], window_bits=8)           # no q: a KZG-style key
    mkey = eng.load_key(bases[:8], kind=2, window_bits=8)
    s = orc.rand_fr_buf(rng, 8).reshape(1, 8, 32)
    out = np.zeros((1, 64), dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    # unknown key id, null pointers
    assert L.vkzg_commit_batch(eng._ctx, ctypes.c_uint32(9999), p(s), ctypes.c_uint32(8), ctypes.c_uint64(1), p(out)) == -2
    assert L.vkzg_commit_batch(eng._ctx, ctypes.c_uint32(wkey.id), None, ctypes.c_uint32(8), ctypes.c_uint64(1), p(out)) == -2
    assert L.vkzg_commit_batch(None, ctypes.c_uint32(wkey.id), p(s), ctypes.c_uint32(8), ctypes.c_uint64(1), p(out)) == -2
    # wrong key kind for the entry point
    with pytest.raises(VkzgError) as e1:
        eng.commit_batch(mkey, s)
    assert e1.value.status == -2
    with pytest.raises(VkzgError) as e2:
        eng.msm(wkey, s[0])
    assert e2.value.status == -2
    # IPA needs q: a key without it refuses to prove / verify
    C = eng.commit_batch(kkey, s)
    with pytest.raises(VkzgError):
        eng.ipa_prove_batch(kkey, s, orc.fr_to_buf([1]), C)
    # width 0 and width beyond the key
    assert L.vkzg_commit_batch(eng._ctx, ctypes.c_uint32(wkey.id), p(s), ctypes.c_uint32(0), ctypes.c_uint64(1), p(out)) == -3
    assert L.vkzg_commit_batch(eng._ctx, ctypes.c_uint32(wkey.id), p(s), ctypes.c_uint32(9), ctypes.c_uint64(1), p(out)) == -3
    # bad window width / key kind at load
    kid = ctypes.c_uint32(0)
    assert L.vkzg_key_load(eng._ctx, p(bases), ctypes.c_uint32(8), None, ctypes.c_uint32(1), ctypes.c_uint32(21), ctypes.byref(kid)) == -2
    assert L.vkzg_key_load(eng._ctx, p(bases), ctypes.c_uint32(8), None, ctypes.c_uint32(7), ctypes.c_uint32(0), ctypes.byref(kid)) == -2
    assert L.vkzg_key_load(eng._ctx, p(bases), ctypes.c_uint32(8), p(bases[8:]), ctypes.c_uint32(2), ctypes.c_uint32(0), ctypes.byref(kid)) == -2
    # multiproof: z outside the key, empty query list
    f = orc.rand_fr_buf(rng, 2 * 8).reshape(2, 8, 32)
    Cq = eng.commit_batch(wkey, f)
    with pytest.raises(VkzgError) as e3:
        eng.multiproof_prove(wkey, "ipa", f, Cq, np.array([0, 8], dtype=np.uint64), f[:, 0])
    assert e3.value.status == -3
    # everything still works after the errors
    assert (eng.commit_batch(wkey, s) == orc.commit_batch(bases[:8], s)).all()
    for k in (wkey, kkey, mkey):
        k.free()
    with pytest.raises(VkzgError):
        eng.commit_batch(wkey, s)                             # freed key


def test_key_load_validation_and_range_checks(eng):
    """off-curve / non-canonical bases are refused at key load (one of them would corrupt unrelated table rows
    through the shared inversions); window widths whose entry lists do not fit shared memory are refused at load, not at the
    first launch; a slice [first, first + n) that wraps around 2^64 is refused; the library works afterwards"""
    from verkle_kzg_b200 import VkzgError, _lib
    L = _lib.lib()
    rng = np.random.default_rng(2)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 9)
    bad = bases.copy()
    bad[3, 32] ^= 1                                            # y changed: not on the curve
    with pytest.raises(VkzgError) as e:
        eng.load_key(bad[:8], window_bits=8)
    assert e.value.status == -2
    with pytest.raises(VkzgError):
        eng.load_key(bad[:8], kind=2, window_bits=8)
    with pytest.raises(VkzgError):
        eng.load_key(bases[:8], q=bad[3], window_bits=8)      # the check covers Q
    noncanon = bases.copy()
    noncanon[0, :32] = np.frombuffer((orc.P_MOD + 5).to_bytes(32, "little"), dtype=np.uint8)   # x >= p
    with pytest.raises(VkzgError):
        eng.load_key(noncanon[:8], window_bits=8)
    with pytest.raises(VkzgError) as e2:
        eng.load_key(bases[:8], window_bits=2)                 # 128 windows: the per-warp entry list exceeds shared memory
    assert e2.value.status == -2
    k3 = eng.load_key(bases[:8], window_bits=3)                # the narrowest that fits
    s = orc.rand_fr_buf(rng, 8).reshape(1, 8, 32)
    assert (eng.commit_batch(k3, s) == orc.commit_batch(bases[:8], s)).all()
    k3.free()
    mk = eng.load_key(bases[:8], kind=2, window_bits=8)
    import torch
    ds = torch.from_numpy(s[0]).cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for first, n in ((2 ** 64 - 1, 2), (9, 0), (4, 5), (2 ** 63, 2 ** 63)):
        st = L.vkzg_msm_range_dev(eng._ctx, ctypes.c_uint32(mk.id), ctypes.c_uint64(first), ctypes.c_void_p(ds.data_ptr()), ctypes.c_uint64(n),
                                  ctypes.c_void_p(out.data_ptr()))
        assert st == -3, (first, n, st)
    assert L.vkzg_msm_range_dev(eng._ctx, ctypes.c_uint32(mk.id), ctypes.c_uint64(4), ctypes.c_void_p(ds.data_ptr()), ctypes.c_uint64(4),
                                ctypes.c_void_p(out.data_ptr())) == 0
    eng.sync()
    assert (out.cpu().numpy() == orc.msm(bases[4:8], s[0, :4])).all()
    mk.free()
    eng.trim()                                                 # cached scratch back to the driver; the context keeps working
    k8 = eng.load_key(bases[:8], window_bits=8)
    assert (eng.commit_batch(k8, s) == orc.commit_batch(bases[:8], s)).all()
    k8.free()


def test_setup_entry_points_reject_bad_arguments(eng):
    """vkzg_ipa_crs_generate[_at], vkzg_kzg_setup_from_secret, vkzg_kzg_powers: status codes, no crashes"""
    from verkle_kzg_b200 import _lib
    L = _lib.lib()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    out = np.zeros((4, 64), dtype=np.uint8)
    nxt = ctypes.c_uint64(0)
    ok = ctypes.c_int32(7)
    seed = ctypes.c_char_p(b"seed")
    # null output / zero count / null seed with a length / null ctx
    assert L.vkzg_ipa_crs_generate(eng._ctx, seed, ctypes.c_uint64(4), ctypes.c_uint64(4), None, ctypes.byref(nxt)) == -2
    assert L.vkzg_ipa_crs_generate(eng._ctx, seed, ctypes.c_uint64(4), ctypes.c_uint64(0), p(out), ctypes.byref(nxt)) == -2
    assert L.vkzg_ipa_crs_generate(eng._ctx, None, ctypes.c_uint64(4), ctypes.c_uint64(4), p(out), ctypes.byref(nxt)) == -2
    assert L.vkzg_ipa_crs_generate(None, seed, ctypes.c_uint64(4), ctypes.c_uint64(4), p(out), ctypes.byref(nxt)) == -2
    assert L.vkzg_ipa_crs_generate(eng._ctx, seed, ctypes.c_uint64(4), ctypes.c_uint64((1 << 28) + 1), p(out), ctypes.byref(nxt)) == -3
    # an empty seed is a valid seed (NULL pointer, length 0); next_index may be NULL
    assert L.vkzg_ipa_crs_generate(eng._ctx, None, ctypes.c_uint64(0), ctypes.c_uint64(4), p(out), None) == 0
    assert (out == orc.ipa_crs_gen(b"", 4)[0]).all()
    assert L.vkzg_ipa_crs_generate_at(eng._ctx, seed, ctypes.c_uint64(4), ctypes.c_uint64(0), p(out), None) == -2
    assert L.vkzg_ipa_crs_generate_at(eng._ctx, seed, ctypes.c_uint64(4), ctypes.c_uint64(0), None, ctypes.byref(ok)) == -2
    # an index far beyond any `max` still answers (the bound is the caller's, PointGeneratorError::OutOfBounds)
    assert L.vkzg_ipa_crs_generate_at(eng._ctx, seed, ctypes.c_uint64(4), ctypes.c_uint64((1 << 63) + 12345), p(out), ctypes.byref(ok)) == 0
    ref = orc.ipa_crs_gen_at(b"seed", (1 << 63) + 12345)
    assert (ok.value == 1) == (ref is not None) and (ref is None or (out[0] == ref).all())
    # KZG setup from the secret: needs a window key whose base 0 is G, a secret, m >= 1
    g = orc.g1_generator()
    wkey = eng.load_key(g[None], window_bits=8)
    mkey = eng.load_key(g[None], kind=2, window_bits=8)
    tau = orc.fr_to_buf([5])[0]
    lag = np.zeros((4, 64), dtype=np.uint8)
    assert L.vkzg_kzg_setup_from_secret(eng._ctx, ctypes.c_uint32(mkey.id), p(tau), ctypes.c_uint32(4), p(lag)) == -2
    assert L.vkzg_kzg_setup_from_secret(eng._ctx, ctypes.c_uint32(wkey.id), None, ctypes.c_uint32(4), p(lag)) == -2
    assert L.vkzg_kzg_setup_from_secret(eng._ctx, ctypes.c_uint32(wkey.id), p(tau), ctypes.c_uint32(0), p(lag)) == -2
    assert L.vkzg_kzg_setup_from_secret(eng._ctx, ctypes.c_uint32(wkey.id), p(tau), ctypes.c_uint32((1 << 24) + 1), p(lag)) == -3
    assert L.vkzg_kzg_setup_from_secret(eng._ctx, ctypes.c_uint32(wkey.id), p(tau), ctypes.c_uint32(4), p(lag)) == 0
    assert (lag == orc.kzg_setup(4, 5)).all()
    wkey.free()
    mkey.free()
