"""Product arithmetic headers (csrc/field.cuh, curve.cuh, hash.cuh) executed on the CPU through
tests/host/libhostcheck.so and compared with Python ints / hashlib / the oracle.  The carry-chain
primitives are host-emulated; everything above them is the code the kernels run."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

import orc
import pyref

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host")


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(_DIR, "libhostcheck.so")
    subprocess.check_call(["make", "-C", _DIR, "libhostcheck.so"], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _op(hc, tag, op, a, b=None):
    out = np.zeros_like(a)
    hc.hc_field_op(tag, op, _p(a), None if b is None else _p(b), _p(out), ctypes.c_uint64(len(a)))
    return out


@pytest.mark.parametrize("tag,mod", [(0, orc.R_MOD), (1, orc.P_MOD)])
def test_field_ops(hc, tag, mod):
    rng = np.random.default_rng(100 + tag)
    enc, dec = (orc.fr_to_buf, orc.buf_to_fr) if tag == 0 else (orc.fq_to_buf, orc.buf_to_fq)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 2 ** 253, 2 ** 32 - 1, 2 ** 224 - 1]
    xs = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(400)] + edge + edge
    ys = [int.from_bytes(rng.bytes(32), "little") % mod for _ in range(400)] + edge + edge[::-1]
    a, b = enc(xs), enc(ys)
    assert dec(_op(hc, tag, 0, a, b)) == [(x + y) % mod for x, y in zip(xs, ys)]
    assert dec(_op(hc, tag, 1, a, b)) == [(x - y) % mod for x, y in zip(xs, ys)]
    assert dec(_op(hc, tag, 2, a, b)) == [(x * y) % mod for x, y in zip(xs, ys)]
    assert dec(_op(hc, tag, 6, a)) == [(-x) % mod for x in xs]
    nz = [x for x in xs[:400] + edge + [3, 5, (mod - 1) // 3, 1 << 200, (1 << 253) + 12345] if x]
    # Montgomery representatives with long runs of trailing zeros (the binary inversion strips them up to 31 at a time)
    rinv0 = pow(orc.MONT_R, -1, mod)
    nz += [((1 << j) * rinv0) % mod for j in (1, 31, 32, 33, 63, 64, 95, 96, 200, 253)]
    nz += [(((1 << j) * 0xdeadbeef1) * rinv0) % mod for j in (31, 32, 64, 128)]
    assert dec(_op(hc, tag, 3, enc(nz))) == [pow(x, -1, mod) for x in nz]
    assert dec(_op(hc, tag, 7, enc(nz))) == [pow(x, -1, mod) for x in nz]
    assert dec(_op(hc, tag, 3, enc([0]))) == [0]
    # from_mont gives the canonical integer; to_mont inverts it
    canon = _op(hc, tag, 4, a)
    assert [int.from_bytes(bytes(r), "little") for r in canon] == xs
    assert (_op(hc, tag, 5, canon) == a).all()
    # worst-case limbs for the carry chains: raw Montgomery representations with all-ones limbs (< p)
    raw = np.full((4, 32), 0xFF, dtype=np.uint8)
    raw[:, 31] = [0x2F, 0x30, 0x1F, 0x00]
    raw[1, 28:31] = [0, 0, 0]
    ints = [int.from_bytes(bytes(r), "little") for r in raw]
    assert all(v < mod for v in ints)
    rinv = pow(orc.MONT_R, -1, mod)
    got = _op(hc, tag, 2, raw, raw[::-1].copy())
    want = [(x * y * rinv) % mod for x, y in zip(ints, ints[::-1])]
    assert [int.from_bytes(bytes(r), "little") for r in got] == want


@pytest.mark.parametrize("tag,mod", [(0, orc.R_MOD), (1, orc.P_MOD)])
def test_lazy_reduction_ops(hc, tag, mod):
    """the hot loops' "almost Montgomery" arithmetic: raw representatives anywhere in [0, 2p), results correct mod p"""
    rng = np.random.default_rng(300 + tag)
    edge = [0, 1, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod - 2, (1 << 254) - 1, (1 << 254), mod + (1 << 253)]
    edge = [e for e in edge if e < 2 * mod]
    xs = [int.from_bytes(rng.bytes(32), "little") % (2 * mod) for _ in range(300)] + edge + edge
    ys = [int.from_bytes(rng.bytes(32), "little") % (2 * mod) for _ in range(300)] + edge + edge[::-1]
    raw = lambda v: np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in v), dtype=np.uint8).reshape(-1, 32).copy()
    a, b = raw(xs), raw(ys)
    rinv = pow(orc.MONT_R, -1, mod)
    val = lambda buf: [int.from_bytes(bytes(r), "little") for r in buf]
    assert val(_op(hc, tag, 8, a, b)) == [(x * y * rinv) % mod for x, y in zip(xs, ys)]
    assert val(_op(hc, tag, 9, a, b)) == [(x - y) % mod for x, y in zip(xs, ys)]
    assert val(_op(hc, tag, 10, a, b)) == [(x + y) % mod for x, y in zip(xs, ys)]


@pytest.mark.parametrize("tag,mod", [(0, orc.R_MOD), (1, orc.P_MOD)])
def test_karatsuba_product(hc, tag, mod):
    """fp_mul_lazy_kara: three 4 x 4-limb products glued on the ALU pipe, then the reduction rows alone; operands anywhere in
    [0, 2p], raw result below 2p.  The operand pairs stress both signs of (a1 - a0)(b1 - b0), equal halves and all-ones limbs."""
    rng = np.random.default_rng(900 + tag)
    half = lambda lo, hi: (hi << 128) | lo
    M128 = (1 << 128) - 1
    edge = [0, 1, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod, (1 << 254) - 1, (1 << 254), mod + (1 << 253), 2 * mod - (1 << 32), (1 << 128), M128,
            half(M128, 0), half(0, (2 * mod) >> 128), half(M128, (2 * mod) >> 128) if half(M128, (2 * mod) >> 128) <= 2 * mod else 0,
            half(12345, 12345), half(M128, M128 >> 3), half(M128 >> 3, M128 >> 3), half(1, 0), half(0, 1), half(1 << 127, 1 << 125)]
    edge = [e for e in edge if e <= 2 * mod]
    xs, ys = [], []
    for e in edge:
        for f in edge:
            xs.append(e)
            ys.append(f)
    for _ in range(4000):
        xs.append(int.from_bytes(rng.bytes(32), "little") % (2 * mod + 1))
        ys.append(int.from_bytes(rng.bytes(32), "little") % (2 * mod + 1))
    for _ in range(500):   # halves that differ only in their low limbs: tiny |a1 - a0|, both signs
        h = int.from_bytes(rng.bytes(16), "little") >> 3
        d1, d2 = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        xs.append(half(h, max(h - d1, 0)))
        ys.append(half(max(h - d2, 0), h))
    raw = lambda v: np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in v), dtype=np.uint8).reshape(-1, 32).copy()
    a, b = raw(xs), raw(ys)
    rinv = pow(orc.MONT_R, -1, mod)
    val = lambda buf: [int.from_bytes(bytes(r), "little") for r in buf]
    want = [(x * y * rinv) % mod for x, y in zip(xs, ys)]
    assert val(_op(hc, tag, 13, a, b)) == want
    got = val(_op(hc, tag, 14, a, b))
    assert all(g < 2 * mod for g in got)
    assert [g % mod for g in got] == want


@pytest.mark.parametrize("tag,mod", [(0, orc.R_MOD), (1, orc.P_MOD)])
def test_dedicated_square(hc, tag, mod):
    """fp_sqr_lazy: a^2 / R with 36 instead of 64 limb products (doubled off-diagonal rows interleaved with the reduction),
    operand anywhere in [0, 2p]; the raw result must respect the [0, 2p) invariant of the hot loops"""
    rng = np.random.default_rng(700 + tag)
    edge = [0, 1, 2, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod, (1 << 254) - 1, (1 << 254), mod + (1 << 253), 2 * mod - (1 << 32),
            (1 << 255) % (2 * mod), (1 << 31), (1 << 32) - 1, (1 << 63), ((1 << 224) - 1) << 29]
    # limb patterns that stress the doubled rows: bit 31 of every limb set / only the top bits / alternating all-ones limbs
    pat = [sum(0x80000000 << (32 * i) for i in range(8)), sum(0xFFFFFFFF << (32 * i) for i in range(0, 8, 2)),
           sum(0xFFFFFFFF << (32 * i) for i in range(1, 8, 2)), sum(0x80000001 << (32 * i) for i in range(8)), (1 << 256) - 1]
    edge += [v % (2 * mod + 1) for v in pat] + [min(v, 2 * mod) for v in pat]
    for top in range(0, 0x61):  # every value of the top byte that keeps the operand <= 2p, all other bits set
        v = (top << 248) | ((1 << 248) - 1)
        if v <= 2 * mod:
            edge.append(v)
    xs = [int.from_bytes(rng.bytes(32), "little") % (2 * mod + 1) for _ in range(3000)] + edge
    raw = lambda v: np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in v), dtype=np.uint8).reshape(-1, 32).copy()
    a = raw(xs)
    rinv = pow(orc.MONT_R, -1, mod)
    val = lambda buf: [int.from_bytes(bytes(r), "little") for r in buf]
    want = [(x * x * rinv) % mod for x in xs]
    assert val(_op(hc, tag, 11, a)) == want
    got = val(_op(hc, tag, 12, a))
    assert all(g < 2 * mod for g in got)
    assert [g % mod for g in got] == want
    # and it is the same function as the general lazy product on equal operands
    assert val(_op(hc, tag, 8, a, a)) == want


@pytest.mark.parametrize("tag,mod", [(0, orc.R_MOD), (1, orc.P_MOD)])
def test_fused_pair_of_products(hc, tag, mod):
    """fp_mul2_lazy: (a b + c d) / R in one interleaved Montgomery pass, operands anywhere in [0, 2p] (2p itself is what
    fp_neg_lazy returns for 0); the raw result must respect the [0, 2p) invariant of the hot loops"""
    rng = np.random.default_rng(500 + tag)
    edge = [0, 1, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod, (1 << 254) - 1, (1 << 254), mod + (1 << 253), 2 * mod - (1 << 32), (1 << 255) % (2 * mod)]
    edge = [e for e in edge if e <= 2 * mod]
    cols = []
    for k in range(4):
        v = [int.from_bytes(rng.bytes(32), "little") % (2 * mod + 1) for _ in range(500)]
        v += edge[k:] + edge[:k] + [2 * mod] * 8 + [2 * mod - 1] * 4   # all-maximal rows: the worst case of the running total
        cols.append(v)
    raw = lambda v: np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in v), dtype=np.uint8).reshape(-1, 32).copy()
    bufs = [raw(v) for v in cols]
    rinv = pow(orc.MONT_R, -1, mod)
    val = lambda buf: [int.from_bytes(bytes(r), "little") for r in buf]
    want = [((a * b + c * d) * rinv) % mod for a, b, c, d in zip(*cols)]
    for rawflag in (0, 1):
        out = np.zeros_like(bufs[0])
        hc.hc_mul2(tag, _p(bufs[0]), _p(bufs[1]), _p(bufs[2]), _p(bufs[3]), _p(out), ctypes.c_uint64(len(out)), rawflag)
        got = val(out)
        if rawflag:
            assert all(g < 2 * mod for g in got)
            got = [g % mod for g in got]
        assert got == want
    neg = np.zeros_like(bufs[0])
    xs = [x for x in cols[0] if x < 2 * mod]
    hc.hc_neg_lazy(tag, _p(raw(xs)), _p(neg), ctypes.c_uint64(len(xs)))
    assert val(neg[:len(xs)]) == [2 * mod - x for x in xs]


def test_group_ops(hc):
    rng = np.random.default_rng(7)
    ks = orc.rand_fr(rng, 6) + [1, 2]
    pts = [pyref.g_mul(pyref.G1_GEN, k) for k in ks] + [None]
    buf = orc.pts_to_buf(pts)
    out = np.zeros(64, dtype=np.uint8)
    for i in range(len(pts)):
        for j in range(len(pts)):
            for mode in (0, 1):
                hc.hc_g1_op(mode, _p(buf[i]), _p(buf[j]), _p(out))
                assert orc.buf_to_pts(out)[0] == pyref.g_add(pts[i], pts[j]), (mode, i, j)
        hc.hc_g1_op(2, _p(buf[i]), _p(buf[i]), _p(out))
        assert orc.buf_to_pts(out)[0] == pyref.g_add(pts[i], pts[i])
        # P + (-P) -> infinity, then + again
        neg = orc.pts_to_buf([pyref.g_neg(pts[i])])[0]
        hc.hc_g1_op(0, _p(buf[i]), _p(neg), _p(out))
        assert orc.buf_to_pts(out)[0] is None
        hc.hc_g1_op(3, _p(buf[i]), _p(buf[(i + 1) % len(pts)]), _p(out))
        assert orc.buf_to_pts(out)[0] == pyref.g_add(pts[i], pts[(i + 1) % len(pts)])
        # the kernels' lazily reduced mixed addition (xyzz_madd_hot + xyzz_canon)
        hc.hc_g1_op(4, _p(buf[i]), _p(buf[(i + 1) % len(pts)]), _p(out))
        assert orc.buf_to_pts(out)[0] == pyref.g_add(pts[i], pts[(i + 1) % len(pts)])
        hc.hc_g1_op(5, _p(buf[i]), _p(buf[(i + 1) % len(pts)]), _p(out))
        assert orc.buf_to_pts(out)[0] == pts[(i + 1) % len(pts)]


@pytest.mark.parametrize("c", [4, 8, 13, 16, 20])
def test_signed_window_recoding_msm(hc, c):
    rng = np.random.default_rng(200 + c)
    n = 6
    bases = orc.points_walk(*orc.rand_fr(rng, 2), n)
    s = orc.rand_fr(rng, n)
    s[0], s[1], s[2] = orc.R_MOD - 1, 0, (1 << (c - 1))
    s[3] = int("1" * 254, 2) % orc.R_MOD
    sb = orc.fr_to_buf(s)
    out = np.zeros(64, dtype=np.uint8)
    hc.hc_msm_windowed(_p(bases), _p(sb), ctypes.c_uint64(n), c, _p(out))
    assert (out == orc.msm(bases, sb, "naive", 2)).all()


def test_compress_and_hashing(hc):
    rng = np.random.default_rng(9)
    pts = [pyref.g_mul(pyref.G1_GEN, k) for k in orc.rand_fr(rng, 8)] + [None, pyref.G1_GEN, pyref.g_neg(pyref.G1_GEN)]
    buf = orc.pts_to_buf(pts)
    out = np.zeros((len(pts), 32), dtype=np.uint8)
    hc.hc_compress(_p(buf), ctypes.c_uint64(len(pts)), _p(out))
    assert [bytes(r) for r in out] == [pyref.ser_g1(p) for p in pts]
    d = np.zeros(32, dtype=np.uint8)
    for n in [0, 1, 55, 56, 63, 64, 65, 100, 119, 120, 128, 187, 249]:
        msg = np.frombuffer(rng.bytes(max(n, 1)), dtype=np.uint8).copy()
        hc.hc_sha256(_p(msg), n, _p(d))
        assert bytes(d) == hashlib.sha256(bytes(msg[:n])).digest()
        for dst in [b"ipa", b"multiproof"]:
            dd = np.frombuffer(dst, dtype=np.uint8).copy()
            u = np.zeros(48, dtype=np.uint8)
            hc.hc_xmd48(_p(msg), n, _p(dd), len(dst), 48, _p(u))
            assert bytes(u) == pyref.expand_message_xmd(bytes(msg[:n]), dst, 48, 48)
            f = np.zeros(32, dtype=np.uint8)
            hc.hc_hash_to_fr(_p(msg), n, _p(dd), len(dst), _p(f))
            assert orc.buf_to_fr(f)[0] == pyref.hash_to_fr(bytes(msg[:n]), dst)
    # the 48-byte expansion also matches the model at the RFC's z_pad = 64
    dd = np.frombuffer(b"QUUX-V01-CS02-with-expander-SHA256-128", dtype=np.uint8).copy()
    u = np.zeros(48, dtype=np.uint8)
    m = np.frombuffer(b"abc", dtype=np.uint8).copy()
    hc.hc_xmd48(_p(m), 3, _p(dd), len(dd), 64, _p(u))
    assert bytes(u) == pyref.expand_message_xmd(b"abc", bytes(dd), 48, 64)
    # from_le_bytes_mod_order on 32 bytes with flag bits set (to_data_item, quirk Q7)
    for _ in range(8):
        raw = np.frombuffer(rng.bytes(32), dtype=np.uint8).copy()
        f = np.zeros(32, dtype=np.uint8)
        hc.hc_fr_from_le32(_p(raw), _p(f))
        assert orc.buf_to_fr(f)[0] == int.from_bytes(bytes(raw), "little") % orc.R_MOD


def test_transcript_walk(hc):
    rng = np.random.default_rng(10)
    ks = orc.rand_fr(rng, 3)
    C, L, R = [pyref.g_mul(pyref.G1_GEN, k) for k in ks]
    z, y = orc.rand_fr(rng, 2)
    for prefix, dst in [(b"", "ipa"), (bytes(range(66)), "multiproof")]:
        tr = pyref.Transcript(dst, prefix)
        tr.append_g(C, "C")
        tr.append_f(z, "input point")
        tr.append_f(y, "output point")
        w = tr.digest("w")
        tr.append_g(L, "L")
        tr.append_g(R, "R")
        x = tr.digest("x")
        wo, xo = np.zeros(32, dtype=np.uint8), np.zeros(32, dtype=np.uint8)
        pre = np.frombuffer(prefix, dtype=np.uint8).copy() if prefix else np.zeros(1, dtype=np.uint8)
        hc.hc_transcript_ipa(_p(pre), len(prefix), dst.encode(), _p(orc.pts_to_buf([C])), _p(orc.fr_to_buf([z])), _p(orc.fr_to_buf([y])),
                             _p(orc.pts_to_buf([L])), _p(orc.pts_to_buf([R])), _p(wo), _p(xo))
        assert orc.buf_to_fr(wo)[0] == w and orc.buf_to_fr(xo)[0] == x


def test_pipeline_pieces_cover_the_batch(hc):
    """vk_common.cuh: pipeline_piece — the pieces of a pipelined upload are never empty, cover the batch exactly, start small
    (B/16, 3B/16) and finish in one launch; batches below 8192 rows are one piece"""
    out = (ctypes.c_uint64 * 64)()
    hc.hc_pipeline_pieces.argtypes = [ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
    for B in [1, 2, 4095, 8191, 8192, 8193, 8207, 9000, 16384, 16385, 65536, 1_000_003]:
        n = hc.hc_pipeline_pieces(B, out, 64)
        pieces = [int(out[i]) for i in range(n)]
        assert n >= 1 and all(p > 0 for p in pieces) and sum(pieces) == B, (B, pieces)
        if B < 8192:
            assert pieces == [B]
        else:
            assert pieces == [B // 16, 3 * B // 16, B - B // 16 - 3 * B // 16], (B, pieces)
