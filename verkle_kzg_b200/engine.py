"""Engine: one vkzg_ctx (one GPU, one stream) with numpy (host) and torch (device) entry points.

Buffers use the ABI layouts: Fr = 32 B little-endian Montgomery (uint8[..., 32]); G1 affine = x || y
(uint8[..., 64], all-zero = infinity).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import KEY_MSM, KEY_WINDOW, check, dptr, hptr, u8


def _log2(n):
    lg = 0
    while (1 << lg) < n:
        lg += 1
    return lg


class Key:
    def __init__(self, engine, key_id, n, has_q, kind):
        self.engine, self.id, self.n, self.has_q, self.kind = engine, key_id, n, has_q, kind
        self.log2n = _log2(n)

    @property
    def table_bytes(self):
        return _lib.lib().vkzg_key_table_bytes(self.engine._ctx, ctypes.c_uint32(self.id))

    def free(self):
        if self.id:
            _lib.lib().vkzg_key_free(self.engine._ctx, ctypes.c_uint32(self.id))
            self.id = 0


class Engine:
    def __init__(self, device=0, stream=None):
        """stream: a raw cudaStream_t (int), e.g. torch.cuda.current_stream().cuda_stream, or None."""
        self._L = _lib.lib()
        self._ctx = ctypes.c_void_p()
        if stream is None:
            st = self._L.vkzg_ctx_create(ctypes.byref(self._ctx), ctypes.c_int32(device))
        else:
            st = self._L.vkzg_ctx_create_on_stream(ctypes.byref(self._ctx), ctypes.c_int32(device), ctypes.c_void_p(stream))
        check(st, "vkzg_ctx_create")
        self.device = device

    def close(self):
        if self._ctx:
            self._L.vkzg_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """return the cached scratch memory of this context's private pool to the driver"""
        check(self._L.vkzg_ctx_trim(self._ctx), "vkzg_ctx_trim")

    def sync(self):
        check(self._L.vkzg_ctx_sync(self._ctx), "vkzg_ctx_sync")

    @property
    def launches(self):
        return int(self._L.vkzg_ctx_launches(self._ctx))

    OPT_IPA_TWO_STREAMS = 1   # include/vkzg.h
    OPT_TREE_FLATTEN = 2      # 0 automatic, 1 bulk pass, 2 depth-first walk of the dirty paths
    OPT_BATCH_AFFINE = 4      # -1 automatic, 0 off, 1 on: batch-affine tree for big dense batches
    OPT_MULTIPROOF_CHECK_Y = 3  # diagnostic (default 0): verify_multiproof also compares y_proof with g2(t) — see include/vkzg.h

    def set_option(self, option, value):
        check(self._L.vkzg_ctx_set_option(self._ctx, ctypes.c_int32(option), ctypes.c_int32(value)), "vkzg_ctx_set_option")

    def kernel_timing(self, enable=True):
        check(self._L.vkzg_ctx_kernel_timing(self._ctx, ctypes.c_int32(1 if enable else 0)), "vkzg_ctx_kernel_timing")

    def kernel_timing_read(self):
        n = ctypes.c_uint64(0)
        ms = ctypes.c_double(0)
        check(self._L.vkzg_ctx_kernel_timing_read(self._ctx, ctypes.byref(n), ctypes.byref(ms)), "vkzg_ctx_kernel_timing_read")
        return n.value, ms.value

    # ---------------------------------------------------------------- keys
    def load_key(self, bases, q=None, kind=KEY_WINDOW, window_bits=0):
        bases = u8(bases, 64).reshape(-1, 64)
        qq = None if q is None else u8(q, 64).reshape(64)
        kid = ctypes.c_uint32(0)
        st = self._L.vkzg_key_load(self._ctx, hptr(bases), ctypes.c_uint32(len(bases)), hptr(qq), ctypes.c_uint32(kind),
                                   ctypes.c_uint32(window_bits), ctypes.byref(kid))
        check(st, "vkzg_key_load")
        return Key(self, kid.value, len(bases), q is not None, kind)

    def load_key_dev(self, d_bases, n, d_q=None, kind=KEY_WINDOW, window_bits=0):
        kid = ctypes.c_uint32(0)
        st = self._L.vkzg_key_load_dev(self._ctx, dptr(d_bases), ctypes.c_uint32(n), dptr(d_q), ctypes.c_uint32(kind),
                                       ctypes.c_uint32(window_bits), ctypes.byref(kid))
        check(st, "vkzg_key_load_dev")
        return Key(self, kid.value, n, d_q is not None, kind)

    # ---------------------------------------------------------------- M1
    def msm(self, key, scalars):
        s = u8(scalars, 32).reshape(-1, 32)
        out = np.zeros(64, dtype=np.uint8)
        check(self._L.vkzg_msm(self._ctx, ctypes.c_uint32(key.id), hptr(s), ctypes.c_uint64(len(s)), hptr(out)), "vkzg_msm")
        return out

    def msm_dev(self, key, d_scalars, n, d_out, first=0):
        check(self._L.vkzg_msm_range_dev(self._ctx, ctypes.c_uint32(key.id), ctypes.c_uint64(first), dptr(d_scalars),
                                         ctypes.c_uint64(n), dptr(d_out)), "vkzg_msm_range_dev")

    def commit_batch(self, key, scalars):
        s = u8(scalars, 32)
        assert s.ndim == 3, "scalars must be [B, w, 32]"
        B, w = s.shape[0], s.shape[1]
        out = np.zeros((B, 64), dtype=np.uint8)
        check(self._L.vkzg_commit_batch(self._ctx, ctypes.c_uint32(key.id), hptr(s), ctypes.c_uint32(w), ctypes.c_uint64(B),
                                        hptr(out)), "vkzg_commit_batch")
        return out

    def commit_batch_dev(self, key, d_scalars, w, B, d_out):
        check(self._L.vkzg_commit_batch_dev(self._ctx, ctypes.c_uint32(key.id), dptr(d_scalars), ctypes.c_uint32(w),
                                            ctypes.c_uint64(B), dptr(d_out)), "vkzg_commit_batch_dev")

    def g1_sum(self, points):
        p = u8(points, 64).reshape(-1, 64)
        out = np.zeros(64, dtype=np.uint8)
        check(self._L.vkzg_g1_sum(self._ctx, hptr(p), ctypes.c_uint64(len(p)), hptr(out)), "vkzg_g1_sum")
        return out

    def g1_sum_dev(self, d_points, n, d_out):
        check(self._L.vkzg_g1_sum_dev(self._ctx, dptr(d_points), ctypes.c_uint64(n), dptr(d_out)), "vkzg_g1_sum_dev")

    def to_data_item(self, points):
        p = u8(points, 64).reshape(-1, 64)
        out = np.zeros((len(p), 32), dtype=np.uint8)
        check(self._L.vkzg_to_data_item(self._ctx, hptr(p), ctypes.c_uint64(len(p)), hptr(out)), "vkzg_to_data_item")
        return out

    # ---------------------------------------------------------------- L1 / I2
    _VEC_OPS = {"add": 0, "sub": 1, "mul": 2, "scale": 3, "axpy": 4}

    def fr_vector_op(self, op, a, b=None, x=None):
        a = u8(a, 32).reshape(-1, 32)
        bb = None if b is None else u8(b, 32).reshape(-1, 32)
        xx = None if x is None else u8(x, 32).reshape(32)
        out = np.zeros_like(a)
        check(self._L.vkzg_fr_vector_op(self._ctx, ctypes.c_int32(self._VEC_OPS[op]), hptr(a), hptr(bb), hptr(xx), ctypes.c_uint64(len(a)),
                                        hptr(out)), "vkzg_fr_vector_op")
        return out

    def fr_vector_op_dev(self, op, d_a, d_b, x, n, d_out):
        xx = None if x is None else u8(x, 32).reshape(32)
        check(self._L.vkzg_fr_vector_op_dev(self._ctx, ctypes.c_int32(self._VEC_OPS[op]), dptr(d_a), dptr(d_b), hptr(xx), ctypes.c_uint64(n),
                                            dptr(d_out)), "vkzg_fr_vector_op_dev")

    # ---------------------------------------------------------------- B1
    def barycentric_batch(self, key, points):
        pts = u8(points, 32).reshape(-1, 32)
        out = np.zeros((len(pts), key.n, 32), dtype=np.uint8)
        check(self._L.vkzg_barycentric_batch(self._ctx, ctypes.c_uint32(key.id), hptr(pts), ctypes.c_uint64(len(pts)), hptr(out)),
              "vkzg_barycentric_batch")
        return out

    # ---------------------------------------------------------------- I1 / I3 / I4
    @staticmethod
    def _prefix(prefix):
        if not prefix:
            return None, 0
        buf = np.frombuffer(bytes(prefix), dtype=np.uint8).copy()
        return buf, len(buf)

    def ipa_prove_batch(self, key, a, points, commitments, prefix=b"", dst="ipa"):
        """a [B,N,32], points [B,32], commitments [B,64] -> (L [B,lg,64], R [B,lg,64], tip [B,32], y [B,32])"""
        a = u8(a, 32)
        B, N = a.shape[0], a.shape[1]
        assert N == key.n
        points = u8(points, 32).reshape(B, 32)
        commitments = u8(commitments, 64).reshape(B, 64)
        lg = key.log2n
        L = np.zeros((B, lg, 64), dtype=np.uint8)
        R = np.zeros((B, lg, 64), dtype=np.uint8)
        tip = np.zeros((B, 32), dtype=np.uint8)
        y = np.zeros((B, 32), dtype=np.uint8)
        pre, plen = self._prefix(prefix)
        check(self._L.vkzg_ipa_prove_batch(self._ctx, ctypes.c_uint32(key.id), hptr(a), hptr(points), hptr(commitments),
                                           ctypes.c_uint64(B), hptr(pre), ctypes.c_uint32(plen), dst.encode(), hptr(L), hptr(R),
                                           hptr(tip), hptr(y)), "vkzg_ipa_prove_batch")
        return L, R, tip, y

    def ipa_commit_prove_batch(self, key, a, points):
        """commit + open in one call -> (C [B,64], L, R, tip, y)"""
        a = u8(a, 32)
        B, N = a.shape[0], a.shape[1]
        assert N == key.n
        points = u8(points, 32).reshape(B, 32)
        lg = key.log2n
        C = np.zeros((B, 64), dtype=np.uint8)
        L = np.zeros((B, lg, 64), dtype=np.uint8)
        R = np.zeros((B, lg, 64), dtype=np.uint8)
        tip = np.zeros((B, 32), dtype=np.uint8)
        y = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_ipa_commit_prove_batch(self._ctx, ctypes.c_uint32(key.id), hptr(a), hptr(points), ctypes.c_uint64(B), hptr(C),
                                                  hptr(L), hptr(R), hptr(tip), hptr(y)), "vkzg_ipa_commit_prove_batch")
        return C, L, R, tip, y

    def ipa_prove_batch_dev(self, key, d_a, d_points, d_commitments, B, d_L, d_R, d_tip, d_y, prefix=b"", dst="ipa"):
        pre, plen = self._prefix(prefix)
        check(self._L.vkzg_ipa_prove_batch_dev(self._ctx, ctypes.c_uint32(key.id), dptr(d_a), dptr(d_points), dptr(d_commitments),
                                               ctypes.c_uint64(B), hptr(pre), ctypes.c_uint32(plen), dst.encode(), dptr(d_L),
                                               dptr(d_R), dptr(d_tip), dptr(d_y)), "vkzg_ipa_prove_batch_dev")

    def ipa_verify_batch(self, key, points, commitments, L, R, tip, y, prefix=b"", dst="ipa"):
        points = u8(points, 32).reshape(-1, 32)
        B = len(points)
        commitments = u8(commitments, 64).reshape(B, 64)
        L = u8(L, 64).reshape(B, key.log2n, 64)
        R = u8(R, 64).reshape(B, key.log2n, 64)
        tip = u8(tip, 32).reshape(B, 32)
        y = u8(y, 32).reshape(B, 32)
        ok = np.zeros(B, dtype=np.int32)
        pre, plen = self._prefix(prefix)
        check(self._L.vkzg_ipa_verify_batch(self._ctx, ctypes.c_uint32(key.id), hptr(points), hptr(commitments), ctypes.c_uint64(B),
                                            hptr(pre), ctypes.c_uint32(plen), dst.encode(), hptr(L), hptr(R), hptr(tip), hptr(y),
                                            hptr(ok)), "vkzg_ipa_verify_batch")
        return ok.astype(bool)

    def ipa_prove_commitment_batch(self, key, a, commitments):
        a = u8(a, 32)
        B, N = a.shape[0], a.shape[1]
        assert N == key.n
        commitments = u8(commitments, 64).reshape(B, 64)
        lg = key.log2n
        L = np.zeros((B, lg, 64), dtype=np.uint8)
        R = np.zeros((B, lg, 64), dtype=np.uint8)
        tip = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_ipa_prove_commitment_batch(self._ctx, ctypes.c_uint32(key.id), hptr(a), hptr(commitments),
                                                      ctypes.c_uint64(B), hptr(L), hptr(R), hptr(tip)),
              "vkzg_ipa_prove_commitment_batch")
        return L, R, tip

    def ipa_verify_commitment_batch(self, key, commitments, L, R, tip):
        """IPA::verify_commitment_proof (ipa/mod.rs:238-265), batched -> bool [B]"""
        commitments = u8(commitments, 64).reshape(-1, 64)
        B = len(commitments)
        L = u8(L, 64).reshape(B, key.log2n, 64)
        R = u8(R, 64).reshape(B, key.log2n, 64)
        tip = u8(tip, 32).reshape(B, 32)
        ok = np.zeros(B, dtype=np.int32)
        check(self._L.vkzg_ipa_verify_commitment_batch(self._ctx, ctypes.c_uint32(key.id), hptr(commitments), ctypes.c_uint64(B),
                                                       hptr(L), hptr(R), hptr(tip), hptr(ok)), "vkzg_ipa_verify_commitment_batch")
        return ok.astype(bool)

    # ---------------------------------------------------------------- E1 / K1 / K2 / K3
    def evaluate_batch(self, key, f, points, domain_n=0):
        f = u8(f, 32)
        B, ln = f.shape[0], f.shape[1]
        points = u8(points, 32).reshape(B, 32)
        out = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_evaluate_batch(self._ctx, ctypes.c_uint32(key.id), hptr(f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                          hptr(points), ctypes.c_uint64(B), hptr(out)), "vkzg_evaluate_batch")
        return out

    def quotient_batch(self, key, f, points, domain_n=0):
        f = u8(f, 32)
        B, ln = f.shape[0], f.shape[1]
        points = u8(points, 32).reshape(B, 32)
        dn = 1 << _log2(max(ln, domain_n))
        q = np.zeros((B, dn, 32), dtype=np.uint8)
        y = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_quotient_batch(self._ctx, ctypes.c_uint32(key.id), hptr(f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                          hptr(points), ctypes.c_uint64(B), hptr(q), hptr(y)), "vkzg_quotient_batch")
        return q, y

    def quotient_batch_dev(self, key, d_f, ln, d_points, B, d_q, d_y, domain_n=0):
        check(self._L.vkzg_quotient_batch_dev(self._ctx, ctypes.c_uint32(key.id), dptr(d_f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                              dptr(d_points), ctypes.c_uint64(B), dptr(d_q), dptr(d_y)), "vkzg_quotient_batch_dev")

    def evaluate_batch_dev(self, key, d_f, ln, d_points, B, d_y, domain_n=0):
        check(self._L.vkzg_evaluate_batch_dev(self._ctx, ctypes.c_uint32(key.id), dptr(d_f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                              dptr(d_points), ctypes.c_uint64(B), dptr(d_y)), "vkzg_evaluate_batch_dev")

    def kzg_open_batch(self, key, f, points, domain_n=0):
        f = u8(f, 32)
        B, ln = f.shape[0], f.shape[1]
        points = u8(points, 32).reshape(B, 32)
        proof = np.zeros((B, 64), dtype=np.uint8)
        y = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_kzg_open_batch(self._ctx, ctypes.c_uint32(key.id), hptr(f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                          hptr(points), ctypes.c_uint64(B), hptr(proof), hptr(y)), "vkzg_kzg_open_batch")
        return proof, y

    def kzg_commit_open_batch(self, key, f, points, domain_n=0):
        """commit + open the same vectors with one upload of the rows -> (commitments [B,64], proof [B,64], y [B,32])"""
        f = u8(f, 32)
        B, ln = f.shape[0], f.shape[1]
        points = u8(points, 32).reshape(B, 32)
        C = np.zeros((B, 64), dtype=np.uint8)
        proof = np.zeros((B, 64), dtype=np.uint8)
        y = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_kzg_commit_open_batch(self._ctx, ctypes.c_uint32(key.id), hptr(f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                                 hptr(points), ctypes.c_uint64(B), hptr(C), hptr(proof), hptr(y)),
              "vkzg_kzg_commit_open_batch")
        return C, proof, y

    def kzg_prove_all_batch(self, key, f, domain_n=0):
        """KZG::prove_all_points (kzg/mod.rs:200-235): f [B,len,32] -> (proof [B,Dn,64], y [B,Dn,32]), Dn = the data domain size"""
        f = u8(f, 32)
        B, ln = f.shape[0], f.shape[1]
        Dn = 1 << _log2(max(ln, domain_n))
        proof = np.zeros((B, Dn, 64), dtype=np.uint8)
        y = np.zeros((B, Dn, 32), dtype=np.uint8)
        check(self._L.vkzg_kzg_prove_all_batch(self._ctx, ctypes.c_uint32(key.id), hptr(f), ctypes.c_uint32(ln), ctypes.c_uint32(domain_n),
                                               ctypes.c_uint64(B), hptr(proof), hptr(y)), "vkzg_kzg_prove_all_batch")
        return proof, y

    def kzg_open_batch_dev(self, key, d_f, ln, d_points, B, d_proof, d_y, domain_n=0):
        check(self._L.vkzg_kzg_open_batch_dev(self._ctx, ctypes.c_uint32(key.id), dptr(d_f), ctypes.c_uint32(ln),
                                              ctypes.c_uint32(domain_n), dptr(d_points),
                                              ctypes.c_uint64(B), dptr(d_proof), dptr(d_y)), "vkzg_kzg_open_batch_dev")

    # ---------------------------------------------------------------- P1 / P2
    def multiproof_prove(self, key, scheme, f, C, z, y):
        """scheme 'ipa' | 'kzg'; f [m,N,32], C [m,64], z [m] uint64, y [m,32] -> dict(D, L, R, tip, y)"""
        f = u8(f, 32)
        m, N = f.shape[0], f.shape[1]
        assert N == key.n
        C = u8(C, 64).reshape(m, 64)
        y = u8(y, 32).reshape(m, 32)
        z = np.ascontiguousarray(z, dtype=np.uint64).reshape(m)
        lg = key.log2n
        D = np.zeros(64, dtype=np.uint8)
        L = np.zeros((lg, 64), dtype=np.uint8)
        R = np.zeros((lg, 64), dtype=np.uint8)
        tip = np.zeros(32, dtype=np.uint8)
        yo = np.zeros(32, dtype=np.uint8)
        check(self._L.vkzg_multiproof_prove(self._ctx, ctypes.c_uint32(key.id), ctypes.c_int32(0 if scheme == "ipa" else 1), hptr(f),
                                            hptr(C), hptr(z), hptr(y), ctypes.c_uint64(m), hptr(D), hptr(L), hptr(R), hptr(tip),
                                            hptr(yo)), "vkzg_multiproof_prove")
        return dict(D=D, L=L, R=R, tip=tip, y=yo)

    def multiproof_prove_batch(self, key, scheme, f, C, z, y, m_each):
        """K multiproofs in one call: f [sum m, N, 32], C [sum m, 64], z [sum m], y [sum m, 32], m_each [K] -> list of K proof dicts"""
        f = u8(f, 32)
        total, N = f.shape[0], f.shape[1]
        assert N == key.n
        m_each = np.ascontiguousarray(m_each, dtype=np.uint64)
        K = len(m_each)
        assert int(m_each.sum()) == total
        C = u8(C, 64).reshape(total, 64)
        y = u8(y, 32).reshape(total, 32)
        z = np.ascontiguousarray(z, dtype=np.uint64).reshape(total)
        lg = key.log2n
        D = np.zeros((K, 64), dtype=np.uint8)
        L = np.zeros((K, lg, 64), dtype=np.uint8)
        R = np.zeros((K, lg, 64), dtype=np.uint8)
        tip = np.zeros((K, 32), dtype=np.uint8)
        yo = np.zeros((K, 32), dtype=np.uint8)
        Lk = L if scheme == "ipa" else np.zeros((K, 64), dtype=np.uint8)   # KZG: one proof point per multiproof
        check(self._L.vkzg_multiproof_prove_batch(self._ctx, ctypes.c_uint32(key.id), ctypes.c_int32(0 if scheme == "ipa" else 1), hptr(f),
                                                  hptr(C), hptr(z), hptr(y), hptr(m_each), ctypes.c_uint64(K), hptr(D), hptr(Lk), hptr(R),
                                                  hptr(tip), hptr(yo)), "vkzg_multiproof_prove_batch")
        if scheme != "ipa":
            L[:, 0] = Lk
        return [dict(D=D[i], L=L[i], R=R[i], tip=tip[i], y=yo[i]) for i in range(K)]

    def multiproof_verify_ipa(self, key, C, z, y, proof):
        C = u8(C, 64).reshape(-1, 64)
        m = len(C)
        y = u8(y, 32).reshape(m, 32)
        z = np.ascontiguousarray(z, dtype=np.uint64).reshape(m)
        ok = np.zeros(1, dtype=np.int32)
        check(self._L.vkzg_multiproof_verify_ipa(self._ctx, ctypes.c_uint32(key.id), hptr(C), hptr(z), hptr(y), ctypes.c_uint64(m),
                                                 hptr(u8(proof["D"], 64)), hptr(u8(proof["L"], 64)), hptr(u8(proof["R"], 64)),
                                                 hptr(u8(proof["tip"], 32)), hptr(u8(proof["y"], 32)), hptr(ok)),
              "vkzg_multiproof_verify_ipa")
        return bool(ok[0])

    # ---------------------------------------------------------------- T1
    def tree_commit_levels(self, key, levels):
        """levels: list of dict(row_ptr uint32[n+1], slot uint16[t], child int32[t], lit uint8[t,32]), leaves first."""
        n = len(levels)
        rp = [np.ascontiguousarray(l["row_ptr"], dtype=np.uint32) for l in levels]
        sl = [np.ascontiguousarray(l["slot"], dtype=np.uint16) for l in levels]
        ch = [np.ascontiguousarray(l["child"], dtype=np.int32) for l in levels]
        li = [u8(l["lit"], 32).reshape(-1, 32) for l in levels]
        # keep one element at least so that the pointer is valid
        li = [x if len(x) else np.zeros((1, 32), dtype=np.uint8) for x in li]
        sl = [x if len(x) else np.zeros(1, dtype=np.uint16) for x in sl]
        ch = [x if len(x) else np.full(1, -1, dtype=np.int32) for x in ch]
        nodes = (ctypes.c_uint64 * n)(*[len(r) - 1 for r in rp])
        P = ctypes.c_void_p * n
        a_rp = P(*[r.ctypes.data for r in rp])
        a_sl = P(*[x.ctypes.data for x in sl])
        a_ch = P(*[x.ctypes.data for x in ch])
        a_li = P(*[x.ctypes.data for x in li])
        out = np.zeros(64, dtype=np.uint8)
        check(self._L.vkzg_tree_commit_levels(self._ctx, ctypes.c_uint32(key.id), ctypes.c_uint32(n), nodes, a_rp, a_sl, a_ch, a_li,
                                              hptr(out)), "vkzg_tree_commit_levels")
        return out

    def tree_level_dev(self, key, d_row_ptr, n_nodes, d_slot, d_child, d_lit, n_terms, d_nodes, d_out):
        check(self._L.vkzg_tree_level_dev(self._ctx, ctypes.c_uint32(key.id), dptr(d_row_ptr), ctypes.c_uint64(n_nodes),
                                          dptr(d_slot), dptr(d_child), dptr(d_lit), ctypes.c_uint64(n_terms), dptr(d_nodes),
                                          dptr(d_out)), "vkzg_tree_level_dev")

    # ---------------------------------------------------------------- KZG setup (next row 8f-2)
    def kzg_setup(self, powers):
        """powers [m,64] = [tau^i]G -> Lagrange-form SRS [next_pow2(m), 64] (group inverse FFT)"""
        p = u8(powers, 64).reshape(-1, 64)
        n = 1 << _log2(len(p))
        out = np.zeros((n, 64), dtype=np.uint8)
        check(self._L.vkzg_kzg_setup(self._ctx, hptr(p), ctypes.c_uint32(len(p)), hptr(out)), "vkzg_kzg_setup")
        return out

    def kzg_setup_from_secret(self, gen_key, tau, m):
        """KZG::setup(m, KZGRandomPointGenerator::new(tau)) -> Lagrange SRS [next_pow2(m), 64]"""
        n = 1 << _log2(m)
        out = np.zeros((n, 64), dtype=np.uint8)
        check(self._L.vkzg_kzg_setup_from_secret(self._ctx, ctypes.c_uint32(gen_key.id), hptr(u8(tau, 32).reshape(32)), ctypes.c_uint32(m),
                                                 hptr(out)), "vkzg_kzg_setup_from_secret")
        return out

    def kzg_powers(self, gen_key, tau, m):
        out = np.zeros((m, 64), dtype=np.uint8)
        check(self._L.vkzg_kzg_powers(self._ctx, ctypes.c_uint32(gen_key.id), hptr(u8(tau, 32).reshape(32)), ctypes.c_uint32(m), hptr(out)),
              "vkzg_kzg_powers")
        return out

    # ---------------------------------------------------------------- IPA CRS generation (next row 8f-4)
    def ipa_crs_generate(self, seed, num):
        """IPAPointGenerator::gen (ipa_point_generator.rs:51-70): (points [num,64], next unused index)"""
        seed = bytes(seed)
        out = np.zeros((num, 64), dtype=np.uint8)
        nxt = ctypes.c_uint64(0)
        check(self._L.vkzg_ipa_crs_generate(self._ctx, ctypes.c_char_p(seed), ctypes.c_uint64(len(seed)), ctypes.c_uint64(num), hptr(out),
                                            ctypes.byref(nxt)), "vkzg_ipa_crs_generate")
        return out, nxt.value

    def ipa_crs_generate_at(self, seed, index):
        """IPAPointGenerator::gen_at (:72-81): the point, or None (InvalidPoint)"""
        seed = bytes(seed)
        out = np.zeros(64, dtype=np.uint8)
        ok = ctypes.c_int32(0)
        check(self._L.vkzg_ipa_crs_generate_at(self._ctx, ctypes.c_char_p(seed), ctypes.c_uint64(len(seed)), ctypes.c_uint64(index),
                                               hptr(out), ctypes.byref(ok)), "vkzg_ipa_crs_generate_at")
        return out if ok.value else None

    # ---------------------------------------------------------------- probes
    def probe_imad(self, kind, blocks, threads, iters):
        macs = ctypes.c_uint64(0)
        check(self._L.vkzg_probe_imad_dev(self._ctx, ctypes.c_uint32(kind), ctypes.c_uint32(blocks), ctypes.c_uint32(threads),
                                          ctypes.c_uint32(iters), ctypes.byref(macs)), "vkzg_probe_imad_dev")
        return macs.value

    def probe_fq_sqr_dev(self, d_x, n, iters, mode=0):
        check(self._L.vkzg_probe_fq_sqr_dev(self._ctx, dptr(d_x), ctypes.c_uint64(n), ctypes.c_uint32(iters), ctypes.c_uint32(mode)),
              "vkzg_probe_fq_sqr_dev")

    def probe_fq_mul_dev(self, d_x, d_y, n, iters):
        check(self._L.vkzg_probe_fq_mul_dev(self._ctx, dptr(d_x), dptr(d_y), ctypes.c_uint64(n), ctypes.c_uint32(iters)),
              "vkzg_probe_fq_mul_dev")


class MultiEngine:
    """vkzg_mgpu_*: one host process driving several GPUs through the C ABI (include/vkzg.h).  `devices`: list of device ids
    (an id may repeat), or None for every visible device."""

    def __init__(self, devices=None):
        self._L = _lib.lib()
        self._L.vkzg_mgpu_size.restype = ctypes.c_uint32
        self._mg = ctypes.c_void_p()
        if devices is None:
            st = self._L.vkzg_mgpu_create(ctypes.byref(self._mg), None, ctypes.c_uint32(0))
        else:
            ids = np.ascontiguousarray(devices, dtype=np.int32)
            st = self._L.vkzg_mgpu_create(ctypes.byref(self._mg), hptr(ids), ctypes.c_uint32(len(ids)))
        check(st, "vkzg_mgpu_create")

    @property
    def size(self):
        return int(self._L.vkzg_mgpu_size(self._mg))

    def close(self):
        if self._mg:
            self._L.vkzg_mgpu_destroy(self._mg)
            self._mg = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_key(self, bases, q=None, kind=KEY_WINDOW, window_bits=0):
        bases = u8(bases, 64).reshape(-1, 64)
        qq = None if q is None else u8(q, 64).reshape(64)
        kid = ctypes.c_uint32(0)
        check(self._L.vkzg_mgpu_key_load(self._mg, hptr(bases), ctypes.c_uint32(len(bases)), hptr(qq), ctypes.c_uint32(kind),
                                         ctypes.c_uint32(window_bits), ctypes.byref(kid)), "vkzg_mgpu_key_load")
        return kid.value, len(bases)

    def free_key(self, key):
        check(self._L.vkzg_mgpu_key_free(self._mg, ctypes.c_uint32(key[0])), "vkzg_mgpu_key_free")

    def msm(self, key, scalars):
        s = u8(scalars, 32).reshape(-1, 32)
        out = np.zeros(64, dtype=np.uint8)
        check(self._L.vkzg_mgpu_msm(self._mg, ctypes.c_uint32(key[0]), hptr(s), ctypes.c_uint64(len(s)), hptr(out)), "vkzg_mgpu_msm")
        return out

    def commit_batch(self, key, scalars):
        s = u8(scalars, 32)
        B, w = s.shape[0], s.shape[1]
        out = np.zeros((B, 64), dtype=np.uint8)
        check(self._L.vkzg_mgpu_commit_batch(self._mg, ctypes.c_uint32(key[0]), hptr(s), ctypes.c_uint32(w), ctypes.c_uint64(B), hptr(out)),
              "vkzg_mgpu_commit_batch")
        return out

    def ipa_commit_prove_batch(self, key, a, points):
        a = u8(a, 32)
        B, N = a.shape[0], a.shape[1]
        assert N == key[1]
        points = u8(points, 32).reshape(B, 32)
        lg = _log2(N)
        C = np.zeros((B, 64), dtype=np.uint8)
        L = np.zeros((B, lg, 64), dtype=np.uint8)
        R = np.zeros((B, lg, 64), dtype=np.uint8)
        tip = np.zeros((B, 32), dtype=np.uint8)
        y = np.zeros((B, 32), dtype=np.uint8)
        check(self._L.vkzg_mgpu_ipa_commit_prove_batch(self._mg, ctypes.c_uint32(key[0]), hptr(a), hptr(points), ctypes.c_uint64(B), hptr(C),
                                                       hptr(L), hptr(R), hptr(tip), hptr(y)), "vkzg_mgpu_ipa_commit_prove_batch")
        return C, L, R, tip, y
