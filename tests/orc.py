"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Buffers are numpy uint8 arrays in the C-ABI layouts of include/vkzg.h:
  Fr/Fq: 32 B little-endian Montgomery (R = 2^256);  G1 affine: x||y 64 B, all-zero = infinity.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "liborc.so")

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
MONT_R = 1 << 256


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "liborc.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ int <-> buffer conversions (pure python, independent of the oracle)
def fr_to_buf(xs):
    """list of canonical ints -> [n,32] uint8 Montgomery."""
    out = np.zeros((len(xs), 32), dtype=np.uint8)
    for i, x in enumerate(xs):
        out[i] = np.frombuffer(((int(x) % R_MOD) * MONT_R % R_MOD).to_bytes(32, "little"), dtype=np.uint8)
    return out


def buf_to_fr(buf):
    buf = np.ascontiguousarray(buf).reshape(-1, 32)
    rinv = pow(MONT_R, -1, R_MOD)
    return [int.from_bytes(bytes(row), "little") * rinv % R_MOD for row in buf]


def fq_to_buf(xs):
    out = np.zeros((len(xs), 32), dtype=np.uint8)
    for i, x in enumerate(xs):
        out[i] = np.frombuffer(((int(x) % P_MOD) * MONT_R % P_MOD).to_bytes(32, "little"), dtype=np.uint8)
    return out


def buf_to_fq(buf):
    buf = np.ascontiguousarray(buf).reshape(-1, 32)
    rinv = pow(MONT_R, -1, P_MOD)
    return [int.from_bytes(bytes(row), "little") * rinv % P_MOD for row in buf]


def pts_to_buf(pts):
    """list of affine tuples (x, y) or None -> [n,64] uint8."""
    out = np.zeros((len(pts), 64), dtype=np.uint8)
    for i, p in enumerate(pts):
        if p is None:
            continue
        out[i, :32] = fq_to_buf([p[0]])[0]
        out[i, 32:] = fq_to_buf([p[1]])[0]
    return out


def buf_to_pts(buf):
    buf = np.ascontiguousarray(buf).reshape(-1, 64)
    res = []
    for row in buf:
        if not row.any():
            res.append(None)
        else:
            x, y = buf_to_fq(row.reshape(2, 32))
            res.append((x, y))
    return res


def rand_fr(rng, n):
    """n uniform Fr elements as canonical ints (rng: numpy Generator)."""
    out = []
    while len(out) < n:
        raw = rng.bytes(32)
        v = int.from_bytes(raw, "little") >> 2
        if v < R_MOD:
            out.append(v)
    return out


def rand_fr_buf(rng, n):
    """n uniform Fr elements directly as a Montgomery buffer (fast path: a uniform canonical value
    times R is still uniform, so sample the Montgomery representation itself by rejection)."""
    out = np.zeros((n, 32), dtype=np.uint8)
    filled = 0
    mod_be = np.frombuffer(R_MOD.to_bytes(32, "big"), dtype=np.uint8)
    while filled < n:
        m = (n - filled) * 2 + 16
        raw = np.frombuffer(rng.bytes(32 * m), dtype=np.uint8).reshape(m, 32).copy()
        raw[:, 31] &= 0x3F
        be = raw[:, ::-1]
        # lexicographic compare against modulus (big-endian)
        diff = be.astype(np.int16) - mod_be.astype(np.int16)
        nz = diff != 0
        first = np.where(nz.any(axis=1), nz.argmax(axis=1), 31)
        ok = diff[np.arange(m), first] < 0
        good = raw[ok]
        k = min(len(good), n - filled)
        out[filled:filled + k] = good[:k]
        filled += k
    return out


# ------------------------------------------------------------------ wrappers
def field_op(tag, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros_like(a)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32)
    rc = lib().orc_field_op(tag, {"add": 0, "sub": 1, "mul": 2, "inv": 3}[op], _p(a), _p(bb), _p(out), ctypes.c_uint64(len(a)))
    assert rc == 0
    return out


def to_mont(tag, canon):
    canon = np.ascontiguousarray(canon, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros_like(canon)
    assert lib().orc_to_mont(tag, _p(canon), _p(out), ctypes.c_uint64(len(canon))) == 0
    return out


def from_mont(tag, mont):
    mont = np.ascontiguousarray(mont, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros_like(mont)
    assert lib().orc_from_mont(tag, _p(mont), _p(out), ctypes.c_uint64(len(mont))) == 0
    return out


def g1_generator():
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_g1_generator(_p(out))
    return out


def g1_add(a, b):
    out = np.zeros(64, dtype=np.uint8)
    assert lib().orc_g1_add(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out)) == 0
    return out


def g1_neg(a):
    out = np.zeros(64, dtype=np.uint8)
    assert lib().orc_g1_neg(_p(np.ascontiguousarray(a)), _p(out)) == 0
    return out


def g1_sub(a, b):
    return g1_add(a, g1_neg(b))


def g1_mul(p, k):
    out = np.zeros(64, dtype=np.uint8)
    assert lib().orc_g1_mul(_p(np.ascontiguousarray(p)), _p(np.ascontiguousarray(k)), _p(out)) == 0
    return out


def g1_on_curve(p):
    return lib().orc_g1_on_curve(_p(np.ascontiguousarray(p))) == 1


def g1_compress(pts):
    pts = np.ascontiguousarray(pts, dtype=np.uint8).reshape(-1, 64)
    out = np.zeros((len(pts), 32), dtype=np.uint8)
    assert lib().orc_g1_compress(_p(pts), _p(out), ctypes.c_uint64(len(pts))) == 0
    return out


def points_walk(k0, k1, n):
    """P_i = (k0 + i*k1) G as [n,64]; k0,k1 canonical ints."""
    out = np.zeros((n, 64), dtype=np.uint8)
    assert lib().orc_points_walk(_p(fr_to_buf([k0])), _p(fr_to_buf([k1])), ctypes.c_uint64(n), _p(out)) == 0
    return out


def g1_mul_gen_batch(kbuf, nthreads=8):
    kbuf = np.ascontiguousarray(kbuf, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros((len(kbuf), 64), dtype=np.uint8)
    assert lib().orc_g1_mul_gen_batch(_p(kbuf), ctypes.c_uint64(len(kbuf)), nthreads, _p(out)) == 0
    return out


def sha256(msg):
    out = np.zeros(32, dtype=np.uint8)
    m = np.frombuffer(bytes(msg), dtype=np.uint8).copy() if len(msg) else np.zeros(1, dtype=np.uint8)
    lib().orc_sha256(_p(m), ctypes.c_uint64(len(msg)), _p(out))
    return bytes(out)


def expand_message_xmd(msg, dst, n, z_pad_len):
    out = np.zeros(n, dtype=np.uint8)
    m = np.frombuffer(bytes(msg), dtype=np.uint8).copy() if len(msg) else np.zeros(1, dtype=np.uint8)
    d = np.frombuffer(bytes(dst), dtype=np.uint8).copy()
    assert lib().orc_expand_message_xmd(_p(m), ctypes.c_uint64(len(msg)), _p(d), ctypes.c_uint64(len(dst)), ctypes.c_uint64(n),
                                        ctypes.c_uint64(z_pad_len), _p(out)) == 0
    return bytes(out)


def hash_to_fr(msg, dst):
    out = np.zeros(32, dtype=np.uint8)
    m = np.frombuffer(bytes(msg), dtype=np.uint8).copy() if len(msg) else np.zeros(1, dtype=np.uint8)
    assert lib().orc_hash_to_fr(_p(m), ctypes.c_uint64(len(msg)), dst.encode(), _p(out)) == 0
    return out


def domain_gen(n):
    out = np.zeros(32, dtype=np.uint8)
    assert lib().orc_domain_gen(ctypes.c_uint64(n), _p(out)) == 0
    return out


def to_data_item(pts):
    pts = np.ascontiguousarray(pts, dtype=np.uint8).reshape(-1, 64)
    out = np.zeros((len(pts), 32), dtype=np.uint8)
    assert lib().orc_to_data_item(_p(pts), ctypes.c_uint64(len(pts)), _p(out)) == 0
    return out


def msm(bases, scalars, mode="naive", nthreads=8):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1, 32)
    n = min(len(bases), len(scalars))
    out = np.zeros(64, dtype=np.uint8)
    assert lib().orc_msm(_p(bases), _p(scalars), ctypes.c_uint64(n), 0 if mode == "naive" else 1, nthreads, _p(out)) == 0
    return out


def commit_batch(bases, scalars, nthreads=8):
    """scalars [B,w,32] -> [B,64]"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8)
    B, w = scalars.shape[0], scalars.shape[1]
    out = np.zeros((B, 64), dtype=np.uint8)
    assert lib().orc_commit_batch(_p(bases), ctypes.c_uint64(len(bases)), _p(scalars), ctypes.c_uint64(w), ctypes.c_uint64(B),
                                  nthreads, _p(out)) == 0
    return out


def barycentric(N, point_buf):
    out = np.zeros((N, 32), dtype=np.uint8)
    assert lib().orc_barycentric(ctypes.c_uint64(N), _p(np.ascontiguousarray(point_buf)), _p(out)) == 0
    return out


def vanishing(N):
    ev = np.zeros((N, 32), dtype=np.uint8)
    inv = np.zeros((N, 32), dtype=np.uint8)
    assert lib().orc_vanishing(ctypes.c_uint64(N), _p(ev), _p(inv)) == 0
    return ev, inv


def evaluate(N, data, domain_n, point_buf):
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
    out = np.zeros(32, dtype=np.uint8)
    assert lib().orc_evaluate(ctypes.c_uint64(N), _p(data), ctypes.c_uint64(len(data)), ctypes.c_uint64(domain_n),
                              _p(np.ascontiguousarray(point_buf)), _p(out)) == 0
    return out


def divide_by_vanishing(N, data, domain_n, index):
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
    n = 1
    while n < domain_n:
        n <<= 1
    out = np.zeros((n, 32), dtype=np.uint8)
    rc = lib().orc_divide_by_vanishing(ctypes.c_uint64(N), _p(data), ctypes.c_uint64(len(data)), ctypes.c_uint64(domain_n),
                                       ctypes.c_uint64(index), _p(out))
    assert rc == 0
    return out


def divide_by_vanishing_outside(N, data, domain_n, point_buf):
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
    n = 1
    while n < domain_n:
        n <<= 1
    out = np.zeros((n, 32), dtype=np.uint8)
    assert lib().orc_divide_by_vanishing_outside(ctypes.c_uint64(N), _p(data), ctypes.c_uint64(len(data)), ctypes.c_uint64(domain_n),
                                                 _p(np.ascontiguousarray(point_buf)), _p(out)) == 0
    return out


def _log2(N):
    lg = 0
    while (1 << lg) < N:
        lg += 1
    return lg


def ipa_prove(bases, N, a, commitment, point_buf, prefix=b"", dst="ipa"):
    """returns (L[lg,64], R[lg,64], tip[32], y[32])"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
    lg = _log2(N)
    L = np.zeros((lg, 64), dtype=np.uint8)
    R = np.zeros((lg, 64), dtype=np.uint8)
    tip = np.zeros(32, dtype=np.uint8)
    y = np.zeros(32, dtype=np.uint8)
    pre = np.frombuffer(bytes(prefix), dtype=np.uint8).copy() if len(prefix) else None
    rc = lib().orc_ipa_prove(_p(bases), ctypes.c_uint64(N), _p(a), _p(np.ascontiguousarray(commitment)),
                             _p(np.ascontiguousarray(point_buf)), _p(pre), ctypes.c_uint64(len(prefix)), dst.encode(), _p(L), _p(R),
                             _p(tip), _p(y))
    assert rc == 0
    return L, R, tip, y


def ipa_prove_batch(bases, N, a, commitments, points, nthreads=8):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    B = a.shape[0]
    lg = _log2(N)
    L = np.zeros((B, lg, 64), dtype=np.uint8)
    R = np.zeros((B, lg, 64), dtype=np.uint8)
    tip = np.zeros((B, 32), dtype=np.uint8)
    y = np.zeros((B, 32), dtype=np.uint8)
    rc = lib().orc_ipa_prove_batch(_p(bases), ctypes.c_uint64(N), _p(a), _p(np.ascontiguousarray(commitments)),
                                   _p(np.ascontiguousarray(points)), ctypes.c_uint64(B), nthreads, _p(L), _p(R), _p(tip), _p(y))
    assert rc == 0
    return L, R, tip, y


def ipa_verify(bases, N, commitment, point_buf, L, R, tip, y, prefix=b"", dst="ipa"):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    L = np.ascontiguousarray(L, dtype=np.uint8).reshape(-1, 64)
    R = np.ascontiguousarray(R, dtype=np.uint8).reshape(-1, 64)
    pre = np.frombuffer(bytes(prefix), dtype=np.uint8).copy() if len(prefix) else None
    rc = lib().orc_ipa_verify(_p(bases), ctypes.c_uint64(N), _p(np.ascontiguousarray(commitment)), _p(np.ascontiguousarray(point_buf)),
                              _p(pre), ctypes.c_uint64(len(prefix)), dst.encode(), _p(L), _p(R), ctypes.c_uint64(len(L)),
                              _p(np.ascontiguousarray(tip)), _p(np.ascontiguousarray(y)))
    assert rc >= 0
    return rc == 1


def ipa_prove_commitment(bases, N, a, commitment):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
    lg = _log2(len(a))
    L = np.zeros((lg, 64), dtype=np.uint8)
    R = np.zeros((lg, 64), dtype=np.uint8)
    tip = np.zeros(32, dtype=np.uint8)
    rc = lib().orc_ipa_prove_commitment(_p(bases), ctypes.c_uint64(N), _p(a), ctypes.c_uint64(len(a)),
                                        _p(np.ascontiguousarray(commitment)), _p(L), _p(R), _p(tip))
    assert rc == 0
    return L, R, tip


def ipa_verify_commitment(bases, N, commitment, L, R, tip):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    L = np.ascontiguousarray(L, dtype=np.uint8).reshape(-1, 64)
    R = np.ascontiguousarray(R, dtype=np.uint8).reshape(-1, 64)
    rc = lib().orc_ipa_verify_commitment(_p(bases), ctypes.c_uint64(N), _p(np.ascontiguousarray(commitment)), _p(L), _p(R),
                                         ctypes.c_uint64(len(L)), _p(np.ascontiguousarray(tip)))
    assert rc >= 0
    return rc == 1


def ipa_crs_gen(seed, num):
    """IPAPointGenerator::gen -> (points [num,64], next unused index)"""
    seed = bytes(seed)
    out = np.zeros((num, 64), dtype=np.uint8)
    nxt = ctypes.c_uint64(0)
    assert lib().orc_ipa_crs_gen(ctypes.c_char_p(seed), ctypes.c_uint64(len(seed)), ctypes.c_uint64(num), _p(out), ctypes.byref(nxt)) == 0
    return out, nxt.value


def ipa_crs_gen_at(seed, index):
    seed = bytes(seed)
    out = np.zeros(64, dtype=np.uint8)
    rc = lib().orc_ipa_crs_gen_at(ctypes.c_char_p(seed), ctypes.c_uint64(len(seed)), ctypes.c_uint64(index), _p(out))
    assert rc >= 0
    return out if rc == 1 else None


def kzg_setup(max_items, tau):
    n = 1
    while n < max_items:
        n <<= 1
    out = np.zeros((n, 64), dtype=np.uint8)
    assert lib().orc_kzg_setup(ctypes.c_uint64(max_items), _p(fr_to_buf([tau])), _p(out)) == 0
    return out


def kzg_prove(lagrange, data, point_buf):
    """returns (proof[64], y[32], ok)"""
    lagrange = np.ascontiguousarray(lagrange, dtype=np.uint8).reshape(-1, 64)
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
    proof = np.zeros(64, dtype=np.uint8)
    y = np.zeros(32, dtype=np.uint8)
    ok = ctypes.c_int(0)
    rc = lib().orc_kzg_prove(_p(lagrange), ctypes.c_uint64(len(lagrange)), _p(data), ctypes.c_uint64(len(data)),
                             _p(np.ascontiguousarray(point_buf)), _p(proof), _p(y), ctypes.byref(ok))
    assert rc == 0
    return proof, y, ok.value == 1


def kzg_verify_tau(lagrange, tau, commitment, point_buf, proof, y):
    lagrange = np.ascontiguousarray(lagrange, dtype=np.uint8).reshape(-1, 64)
    rc = lib().orc_kzg_verify_tau(_p(lagrange), ctypes.c_uint64(len(lagrange)), _p(fr_to_buf([tau])), _p(np.ascontiguousarray(commitment)),
                                  _p(np.ascontiguousarray(point_buf)), _p(np.ascontiguousarray(proof)), _p(np.ascontiguousarray(y)))
    assert rc >= 0
    return rc == 1


def multiproof_prove(scheme, bases, N, f, C, z, y):
    """scheme 'ipa'|'kzg'.  returns dict(D, L, R, tip, y)"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    f = np.ascontiguousarray(f, dtype=np.uint8)
    C = np.ascontiguousarray(C, dtype=np.uint8)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    m = f.shape[0]
    lg = _log2(N)
    D = np.zeros(64, dtype=np.uint8)
    L = np.zeros((lg, 64), dtype=np.uint8)
    R = np.zeros((lg, 64), dtype=np.uint8)
    tip = np.zeros(32, dtype=np.uint8)
    yo = np.zeros(32, dtype=np.uint8)
    rc = lib().orc_multiproof_prove(0 if scheme == "ipa" else 1, _p(bases), ctypes.c_uint64(N), _p(f), _p(C),
                                    z.ctypes.data_as(ctypes.c_void_p), _p(y), ctypes.c_uint64(m), _p(D), _p(L), _p(R), _p(tip), _p(yo))
    assert rc == 0
    return dict(D=D, L=L, R=R, tip=tip, y=yo)


def multiproof_verify(scheme, bases, N, C, z, y, proof, tau=0):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    C = np.ascontiguousarray(C, dtype=np.uint8)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    L = np.ascontiguousarray(proof["L"], dtype=np.uint8).reshape(-1, 64)
    R = np.ascontiguousarray(proof["R"], dtype=np.uint8).reshape(-1, 64)
    rc = lib().orc_multiproof_verify(0 if scheme == "ipa" else 1, _p(bases), ctypes.c_uint64(N), _p(fr_to_buf([tau])), _p(C),
                                     z.ctypes.data_as(ctypes.c_void_p), _p(y), ctypes.c_uint64(len(z)), _p(np.ascontiguousarray(proof["D"])),
                                     _p(L), _p(R), ctypes.c_uint64(len(L)), _p(np.ascontiguousarray(proof["tip"])),
                                     _p(np.ascontiguousarray(proof["y"])))
    assert rc >= 0
    return rc == 1


def tree_commit(bases, keys, values, ext_width=256):
    """keys [n,key_len] uint8, values [n,32] uint8 -> root commitment [64]"""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 64)
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    values = np.ascontiguousarray(values, dtype=np.uint8)
    out = np.zeros(64, dtype=np.uint8)
    rc = lib().orc_tree_commit(_p(bases), ctypes.c_uint64(len(bases)), _p(keys), ctypes.c_uint64(keys.shape[1]), _p(values),
                               ctypes.c_uint64(len(keys)), ctypes.c_uint64(ext_width), _p(out))
    assert rc == 0, "oracle tree_commit failed (the reference would panic on these keys)"
    return out
