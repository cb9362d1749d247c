"""Pinning against the REAL reference (arkworks 0.4 + SleepingShell/verkle-kzg).

`tests/golden/arkworks/make_vectors.py <checkout>` runs two generator tests inside a scratch copy of the reference and
writes `tests/golden/arkworks/vectors.json`.  When that file is present, `test_oracle_matches_arkworks` (CPU) and
`test_libvkzg_matches_arkworks` (GPU, through the C ABI) compare every section byte for byte; when it is absent they SKIP
with "parity unpinned" (this image has no Rust toolchain — DESIGN.md section 6 lists the three conventions a mismatch
would point at: the compressed-point flag bits, DefaultFieldHasher's Z_pad = 48, the radix-2 domain generator).

Two tests always run: the textual patches of make_vectors.py still apply to the reference checkout (when one is on
disk), and the loader/checker code path is exercised end to end on a vectors document built by the oracle itself (so a
real vectors.json cannot fail for reasons of schema drift)."""
import json
import os

import numpy as np
import pytest

import orc
import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = os.path.join(HERE, "golden", "arkworks", "vectors.json")
UNPINNED = "parity unpinned: tests/golden/arkworks/vectors.json absent (make it with tests/golden/arkworks/make_vectors.py on a box with cargo)"


# ------------------------------------------------------------------ encodings
def fr_hex(buf):
    """Montgomery Fr buffer [32] -> hex of the canonical little-endian bytes (ark-serialize)"""
    return bytes(orc.from_mont(0, np.ascontiguousarray(buf).reshape(1, 32))[0]).hex()


def fr_from_hex(h):
    return orc.fr_to_buf([int.from_bytes(bytes.fromhex(h), "little")])[0]


def pc_hex(pt):
    return bytes(orc.g1_compress(np.ascontiguousarray(pt).reshape(1, 64))[0]).hex()


def pt_from_xy_hex(h):
    """ark-serialize uncompressed x || y (flags in the two top bits of the last byte) -> Montgomery affine [64]"""
    b = bytearray(bytes.fromhex(h))
    flags = b[63] & 0xC0
    b[63] &= 0x3F
    if flags & 0x40:
        return np.zeros(64, dtype=np.uint8)
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little")
    return orc.pts_to_buf([(x, y)])[0]


def xy_hex_noflags(pt):
    p = orc.buf_to_pts(np.ascontiguousarray(pt).reshape(1, 64))[0]
    if p is None:
        return (b"\0" * 64).hex()
    return (p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")).hex()


def strip_flags(h):
    b = bytearray(bytes.fromhex(h))
    b[63] &= 0x3F
    return bytes(b).hex()


def k_times_g(k):
    g = orc.g1_generator()
    if k == 0:
        return np.zeros(64, dtype=np.uint8)
    p = orc.g1_mul(g, orc.fr_to_buf([abs(k)])[0])
    return orc.g1_neg(p) if k < 0 else p


def ramp(n, a, b):
    return orc.fr_to_buf([(a + b * i) % orc.R_MOD for i in range(n)])


def tree_value(seed):
    return np.array([(seed * (i + 1)) % 256 for i in range(32)], dtype=np.uint8)


# ------------------------------------------------------------------ implementations under test
class OracleImpl:
    """the CPU restatement (oracle/)"""
    name = "oracle"

    def to_data_item(self, pts):
        return orc.to_data_item(pts)

    def commit(self, bases, a):
        return orc.commit_batch(bases[:len(a)], a[None])[0]

    def ipa_prove(self, bases, N, a, C, zb):
        return orc.ipa_prove(bases, N, a, C, zb)

    def ipa_prove_commitment(self, bases, N, a, C):
        return orc.ipa_prove_commitment(bases, N, a, C)

    def kzg_setup(self, m, tau):
        return orc.kzg_setup(m, tau)

    def kzg_prove(self, lagrange, data, zb):
        pf, y, ok = orc.kzg_prove(lagrange, data, zb)
        assert ok
        return pf, y

    def multiproof(self, bases, N, f, C, z, y):
        return orc.multiproof_prove("ipa", bases, N, f, C, z, y)

    def tree_root(self, srs, keys, vals, ext_width):
        return orc.tree_commit(srs, keys, vals, ext_width=ext_width)

    def ipa_crs(self, seed, num):
        return orc.ipa_crs_gen(seed, num)[0]


class LibImpl:
    """libvkzg through the C ABI (GPU)"""
    name = "libvkzg"

    def __init__(self, eng):
        self.eng = eng

    def to_data_item(self, pts):
        return self.eng.to_data_item(pts)

    def _key(self, bases, N, q=True, wb=8):
        return self.eng.load_key(bases[:N], q=bases[N] if q else None, window_bits=wb)

    def commit(self, bases, a):
        k = self.eng.load_key(bases[:len(a)], window_bits=8)
        try:
            return self.eng.commit_batch(k, a[None])[0]
        finally:
            k.free()

    def ipa_prove(self, bases, N, a, C, zb):
        k = self._key(bases, N)
        try:
            L, R, tip, y = self.eng.ipa_prove_batch(k, a[None], zb[None], C[None])
            assert self.eng.ipa_verify_batch(k, zb[None], C[None], L, R, tip, y).all()
            return L[0], R[0], tip[0], y[0]
        finally:
            k.free()

    def ipa_prove_commitment(self, bases, N, a, C):
        k = self._key(bases, N)
        try:
            L, R, tip = self.eng.ipa_prove_commitment_batch(k, a[None], C[None])
            assert self.eng.ipa_verify_commitment_batch(k, C[None], L, R, tip).all()
            return L[0], R[0], tip[0]
        finally:
            k.free()

    def ipa_crs(self, seed, num):
        return self.eng.ipa_crs_generate(seed, num)[0]

    def kzg_setup(self, m, tau):
        g = orc.g1_generator()
        gk = self.eng.load_key(g[None], window_bits=8)
        try:
            powers = self.eng.kzg_powers(gk, orc.fr_to_buf([tau])[0], m)
        finally:
            gk.free()
        return self.eng.kzg_setup(powers)

    def kzg_prove(self, lagrange, data, zb):
        k = self.eng.load_key(lagrange, window_bits=8)
        try:
            pf, y = self.eng.kzg_open_batch(k, data[None], zb[None], domain_n=len(lagrange))
            return pf[0], y[0]
        finally:
            k.free()

    def multiproof(self, bases, N, f, C, z, y):
        k = self._key(bases, N)
        try:
            return self.eng.multiproof_prove(k, "ipa", f, C, z, y)
        finally:
            k.free()

    def tree_root(self, srs, keys, vals, ext_width):
        from verkle_kzg_b200.tree import NativeVerkleTree
        k = self.eng.load_key(srs, window_bits=8)
        t = NativeVerkleTree(keys.shape[1], ext_width)
        try:
            t.insert_many(keys, vals)
            return t.commitment(self.eng, k)
        finally:
            t.close()
            k.free()


# ------------------------------------------------------------------ the checker
def check_against(vec, impl, scheme_level=True):
    """compare every section of a vectors document with `impl`; returns the number of byte-exact comparisons made"""
    n = 0
    # 1. compressed points / flags / to_data_item  (transcript.rs:64-71, lib.rs:56-67)
    sec = vec["points"]
    pts = np.stack([k_times_g(k) for k in sec["multiples_of_g"]])
    for i, k in enumerate(sec["multiples_of_g"]):
        assert pc_hex(pts[i]) == sec["compressed"][i], f"compressed {k}G: flag / encoding convention differs from arkworks"
        assert xy_hex_noflags(pts[i]) == strip_flags(sec["xy"][i]), f"{k}G coordinates"
        n += 2
    di = impl.to_data_item(pts)
    for i, k in enumerate(sec["multiples_of_g"]):
        assert fr_hex(di[i]) == sec["to_data_item"][i], f"to_data_item({k}G) [{impl.name}]"
        n += 1
    # 2. DefaultFieldHasher
    for h in vec["hash_to_field"]:
        assert fr_hex(orc.hash_to_fr(h["msg"].encode(), h["dst"])) == h["out"], f"hash_to_field({h['msg']!r}, {h['dst']}): Z_pad / expander convention"
        assert pyref.hash_to_fr(h["msg"].encode(), h["dst"].encode()) == int.from_bytes(bytes.fromhex(h["out"]), "little")
        n += 2
    # 3. TranscriptHasher
    t = vec["transcript"]
    tr = pyref.Transcript("ipa")
    G1 = pyref.G1_GEN
    tr.append_g(pyref.g_mul(G1, t["c"]), "C")
    tr.append_f(t["z"], "input point")
    tr.append_f(t["y"], "output point")
    w = tr.digest("w")
    tr.append_g(pyref.g_mul(G1, t["l"]), "L")
    tr.append_g(pyref.g_mul(G1, t["r"]) if t["r"] else None, "R")
    x = tr.digest("x")
    tr.append_usize(t["usize"], "z")
    u = tr.digest("u", clear=False)
    u2 = tr.digest("u")
    for got, key in ((w, "w"), (x, "x"), (u, "u_noclear"), (u2, "u_clear")):
        assert got == int.from_bytes(bytes.fromhex(t[key]), "little"), f"transcript digest {key}"
        n += 1
    # 4. domain generators
    for d in vec["domains"]:
        assert fr_hex(orc.domain_gen(d["n"])) == d["group_gen"], f"group_gen({d['n']})"
        n += 1
    if not scheme_level:
        return n
    # 5. IPA CRS generation (ipa_point_generator.rs:51-109)
    if "ipa_crs" in vec and hasattr(impl, "ipa_crs"):
        crs = impl.ipa_crs(vec["ipa_crs"]["seed"].encode(), len(vec["ipa_crs"]["xy"]))
        for i, h in enumerate(vec["ipa_crs"]["xy"]):
            assert xy_hex_noflags(crs[i]) == strip_flags(h), f"IPA CRS point {i} [{impl.name}]"
            n += 1
    # 6. IPA
    for case in vec["ipa"]:
        N = case["n"]
        bases = orc.points_walk(1, 1, N + 1)           # g_i = (i + 1) G, q = (N + 1) G
        a = ramp(N, 7, 3)
        C = impl.commit(bases, a)
        assert pc_hex(C) == case["commit"] and xy_hex_noflags(C) == strip_flags(case["commit_xy"]), f"IPA commit N={N} [{impl.name}]"
        assert fr_hex(impl.to_data_item(C[None])[0]) == case["to_data_item"]
        n += 3
        for p in case["proofs"]:
            zb = orc.fr_to_buf([p["index"]])[0]
            L, R, tip, y = impl.ipa_prove(bases, N, a, C, zb)
            assert [pc_hex(v) for v in L] == p["l"] and [pc_hex(v) for v in R] == p["r"], f"IPA proof N={N} index={p['index']} [{impl.name}]"
            assert fr_hex(tip) == p["tip"] and fr_hex(y) == p["y"], f"IPA proof N={N} index={p['index']} [{impl.name}]"
            n += 4
        L, R, tip = impl.ipa_prove_commitment(bases, N, a, C)
        cp = case["commit_proof"]
        assert [pc_hex(v) for v in L] == cp["l"] and [pc_hex(v) for v in R] == cp["r"] and fr_hex(tip) == cp["tip"], f"commit proof N={N} [{impl.name}]"
        n += 3
    # 7. KZG
    kz = vec["kzg"]
    lag = impl.kzg_setup(kz["key"], kz["tau"])
    for i, h in enumerate(kz["lagrange_xy"]):
        assert xy_hex_noflags(lag[i]) == strip_flags(h), f"KZG Lagrange SRS point {i} [{impl.name}]"
        n += 1
    data = ramp(8, 9, 5)
    assert pc_hex(impl.commit(lag, data)) == kz["commit"], f"KZG commit [{impl.name}]"
    for p in kz["proofs"]:
        pf, y = impl.kzg_prove(lag, data, orc.fr_to_buf([p["index"]])[0])
        assert pc_hex(pf) == p["proof"] and fr_hex(y) == p["y"], f"KZG proof index={p['index']} [{impl.name}]"
        n += 2
    # 8. multiproof
    mp = vec["multiproof"]
    N = mp["n"]
    bases = orc.points_walk(1, 1, N + 1)
    f = np.stack([ramp(N, 100 * q + 1, q + 2) for q in range(len(mp["z"]))])
    C = np.stack([impl.commit(bases, f[q]) for q in range(len(f))])
    assert [pc_hex(c) for c in C] == mp["commits"]
    z = np.array(mp["z"], dtype=np.uint64)
    y = f[np.arange(len(z)), z.astype(np.int64)]
    got = impl.multiproof(bases, N, f, C, z, y)
    assert pc_hex(got["D"]) == mp["d"], f"multiproof D [{impl.name}]"
    assert [pc_hex(v) for v in got["L"]] == mp["l"] and [pc_hex(v) for v in got["R"]] == mp["r"], f"multiproof L/R [{impl.name}]"
    assert fr_hex(got["tip"]) == mp["tip"] and fr_hex(got["y"]) == mp["y"], f"multiproof tip/y [{impl.name}]"
    n += 5
    # 9. tree
    if "tree" in vec:
        tv = vec["tree"]
        srs = impl.kzg_setup(256, 100)
        keys = np.array(tv["keys"], dtype=np.uint8)
        vals = np.stack([tree_value(s) for s in tv["value_seed"]])
        for k in range(1, len(keys) + 1):
            root = impl.tree_root(srs, keys[:k], vals[:k], tv["key_len"])
            assert pc_hex(root) == tv["roots_after_each_insert"][k - 1], f"tree root after {k} inserts [{impl.name}]"
            n += 1
    return n


def oracle_made_vectors():
    """a vectors document with arkworks' role played by the oracle: exercises the loader, pins nothing"""
    o = OracleImpl()
    ks = [1, 2, -1, 0, 5, -5]
    pts = np.stack([k_times_g(k) for k in ks])

    def xy_flags(pt):  # ark-serialize puts the point flags into the last byte of an uncompressed point too
        b = bytearray(bytes.fromhex(xy_hex_noflags(pt)))
        b[63] |= orc.g1_compress(pt.reshape(1, 64))[0][31] & 0xC0
        return bytes(b).hex()
    vec = {"points": {"multiples_of_g": ks, "compressed": [pc_hex(p) for p in pts], "xy": [xy_flags(p) for p in pts],
                      "to_data_item": [fr_hex(v) for v in orc.to_data_item(pts)]}}
    vec["hash_to_field"] = [{"dst": d, "msg": m, "out": fr_hex(orc.hash_to_fr(m.encode(), d))} for d, m in (("ipa", "abc"), ("multiproof", ""))]
    tr = pyref.Transcript("ipa")
    G1 = pyref.G1_GEN
    tr.append_g(pyref.g_mul(G1, 5), "C")
    tr.append_f(7, "input point")
    tr.append_f(11, "output point")
    w = tr.digest("w")
    tr.append_g(pyref.g_mul(G1, 2), "L")
    tr.append_g(None, "R")
    x = tr.digest("x")
    tr.append_usize(1234567, "z")
    u = tr.digest("u", clear=False)
    u2 = tr.digest("u")
    le = lambda v: int(v).to_bytes(32, "little").hex()
    vec["transcript"] = {"c": 5, "z": 7, "y": 11, "l": 2, "r": 0, "usize": 1234567, "w": le(w), "x": le(x), "u_noclear": le(u), "u_clear": le(u2)}
    vec["domains"] = [{"n": n, "group_gen": fr_hex(orc.domain_gen(n))} for n in (2, 4, 32, 256)]
    cases = []
    for N in (4,):
        bases = orc.points_walk(1, 1, N + 1)
        a = ramp(N, 7, 3)
        C = o.commit(bases, a)
        proofs = []
        for idx in (1, N - 1, N + 1, 2 * N):
            L, R, tip, y = o.ipa_prove(bases, N, a, C, orc.fr_to_buf([idx])[0])
            proofs.append({"index": idx, "l": [pc_hex(v) for v in L], "r": [pc_hex(v) for v in R], "tip": fr_hex(tip), "y": fr_hex(y)})
        L, R, tip = o.ipa_prove_commitment(bases, N, a, C)
        cases.append({"n": N, "commit": pc_hex(C), "commit_xy": xy_flags(C), "to_data_item": fr_hex(orc.to_data_item(C[None])[0]), "proofs": proofs,
                      "commit_proof": {"l": [pc_hex(v) for v in L], "r": [pc_hex(v) for v in R], "tip": fr_hex(tip)}})
    vec["ipa"] = cases
    lag = o.kzg_setup(16, 100)
    data = ramp(8, 9, 5)
    vec["kzg"] = {"tau": 100, "key": 16, "lagrange_xy": [xy_flags(p) for p in lag], "commit": pc_hex(o.commit(lag, data)), "proofs": []}
    for idx in (0, 3, 9, 17):
        pf, y = o.kzg_prove(lag, data, orc.fr_to_buf([idx])[0])
        vec["kzg"]["proofs"].append({"index": idx, "proof": pc_hex(pf), "y": fr_hex(y)})
    N = 32
    bases = orc.points_walk(1, 1, N + 1)
    zs = [3, 31, 3, 0, 17]
    f = np.stack([ramp(N, 100 * q + 1, q + 2) for q in range(5)])
    C = np.stack([o.commit(bases, f[q]) for q in range(5)])
    z = np.array(zs, dtype=np.uint64)
    got = o.multiproof(bases, N, f, C, z, f[np.arange(5), z.astype(np.int64)])
    vec["multiproof"] = {"n": N, "z": zs, "commits": [pc_hex(c) for c in C], "d": pc_hex(got["D"]), "l": [pc_hex(v) for v in got["L"]],
                         "r": [pc_hex(v) for v in got["R"]], "tip": fr_hex(got["tip"]), "y": fr_hex(got["y"])}
    keys = [[1, 2, 3], [1, 2, 4], [1, 7, 0], [200, 0, 255], [1, 9, 9]]
    seeds = [10, 11, 12, 13, 14]
    srs = o.kzg_setup(256, 100)
    ka, va = np.array(keys, dtype=np.uint8), np.stack([tree_value(s) for s in seeds])
    vec["tree"] = {"key_len": 3, "keys": keys, "value_seed": seeds,
                   "roots_after_each_insert": [pc_hex(o.tree_root(srs, ka[:k], va[:k], 3)) for k in range(1, 6)]}
    return vec


# ------------------------------------------------------------------ tests
def test_loader_runs_end_to_end_on_oracle_made_vectors():
    vec = json.loads(json.dumps(oracle_made_vectors()))        # through JSON, like the real file
    assert check_against(vec, OracleImpl()) > 60


def test_generator_patches_apply_to_the_reference():
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "vector-commit")):
        pytest.skip("no reference checkout on this box")
    import importlib.util
    import shutil
    import tempfile
    spec = importlib.util.spec_from_file_location("make_vectors", os.path.join(HERE, "golden", "arkworks", "make_vectors.py"))
    mv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mv)
    tmp = tempfile.mkdtemp(prefix="arkvec_test_")
    try:
        for crate in ("vector-commit", "verkle-tree"):
            shutil.copytree(os.path.join(ref, crate), os.path.join(tmp, crate))
        mv.patch(tmp)                                           # raises SystemExit when a patch no longer applies
        lib = open(os.path.join(tmp, "vector-commit/src/lib.rs")).read()
        assert "mod arkworks_vectors;" in lib and os.path.exists(os.path.join(tmp, "vector-commit/src/arkworks_vectors.rs"))
        assert "fn arkworks_tree_vector()" in open(os.path.join(tmp, "verkle-tree/src/lib.rs")).read()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_oracle_matches_arkworks():
    if not os.path.exists(VEC):
        pytest.skip(UNPINNED)
    assert check_against(json.load(open(VEC)), OracleImpl()) > 60


@pytest.mark.gpu
def test_libvkzg_matches_arkworks():
    if not os.path.exists(VEC):
        pytest.skip(UNPINNED)
    from verkle_kzg_b200 import Engine
    eng = Engine(0)
    try:
        assert check_against(json.load(open(VEC)), LibImpl(eng)) > 60
    finally:
        eng.close()


@pytest.mark.gpu
def test_libvkzg_passes_the_same_checker_on_oracle_made_vectors():
    """the GPU side of the loader (LibImpl: keys, setup, proofs, tree through the C ABI) against the oracle-made document"""
    from verkle_kzg_b200 import Engine
    eng = Engine(0)
    try:
        assert check_against(json.loads(json.dumps(oracle_made_vectors())), LibImpl(eng)) > 60
    finally:
        eng.close()
