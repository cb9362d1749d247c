// TEST INFRASTRUCTURE: runs the product's __host__ __device__ arithmetic (csrc/*.cuh) on the CPU so the
// exact control flow that the kernels execute can be checked against the oracle / Python model without
// a GPU.  The carry-chain primitives are emulated on the host (see field.cuh); everything above them is
// the same code the device runs.
#include "../../verkle_kzg_b200/csrc/hash.cuh"
#include "../../verkle_kzg_b200/csrc/field_kara.cuh"
#include "../../verkle_kzg_b200/csrc/vk_common.cuh"  // host-side plumbing (pipeline_piece)
#include <cstring>
#include <vector>

using namespace vk;

static fp_t ld(const uint8_t* p) { fp_t r; memcpy(r.l, p, 32); return r; }
static void st(uint8_t* p, const fp_t& v) { memcpy(p, v.l, 32); }
static affine_t lda(const uint8_t* p) { affine_t a; a.x = ld(p); a.y = ld(p + 32); return a; }
static void sta(uint8_t* p, const affine_t& a) { st(p, a.x); st(p + 32, a.y); }

extern "C" {

// the pieces a pipelined batch upload of B rows is cut into (vk_common.cuh); returns their count
int hc_pipeline_pieces(uint64_t B, uint64_t* out, int cap) {
    int n = 0;
    for (uint64_t b0 = 0, nb = 0; b0 < B; b0 += nb) {
        nb = pipeline_piece(B, b0);
        if (n < cap) out[n] = nb;
        ++n;
        if (nb == 0 || n > 64) return -1;
    }
    return n;
}


// tag 0 = Fr, 1 = Fq ; op 0 add 1 sub 2 mul 3 inv (binary Euclid) 4 from_mont 5 to_mont 6 neg 7 inv (Fermat) 8 lazy mul (inputs < 2p) 9 lazy sub 10 lazy add 11 lazy square (input <= 2p) 12 lazy square, raw result (must stay below 2p) 13 / 14 Karatsuba lazy product, canonical / raw — outputs canonicalised
int hc_field_op(int tag, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        fp_t x = ld(a + 32 * i), y = b ? ld(b + 32 * i) : x, r;
        if (tag == 0) {
            r = op == 0 ? fp_add<S>(x, y) : op == 1 ? fp_sub<S>(x, y) : op == 2 ? fp_mul<S>(x, y) : op == 3 ? fp_inv<S>(x)
              : op == 4 ? fp_from_mont<S>(x) : op == 5 ? fp_to_mont<S>(x) : op == 7 ? fp_inv_fermat<S>(x)
              : op == 8 ? fp_canon<S>(fp_mul_lazy<S>(x, y)) : op == 9 ? fp_canon<S>(fp_sub_lazy<S>(x, y)) : op == 10 ? fp_canon<S>(fp_add_lazy<S>(x, y)) : op == 11 ? fp_canon<S>(fp_sqr_lazy<S>(x)) : op == 12 ? fp_sqr_lazy<S>(x) : op == 13 ? fp_canon<S>(fp_mul_lazy_kara<S>(x, y)) : op == 14 ? fp_mul_lazy_kara<S>(x, y) : fp_neg<S>(x);
        } else {
            r = op == 0 ? fp_add<Q>(x, y) : op == 1 ? fp_sub<Q>(x, y) : op == 2 ? fp_mul<Q>(x, y) : op == 3 ? fp_inv<Q>(x)
              : op == 4 ? fp_from_mont<Q>(x) : op == 5 ? fp_to_mont<Q>(x) : op == 7 ? fp_inv_fermat<Q>(x)
              : op == 8 ? fp_canon<Q>(fp_mul_lazy<Q>(x, y)) : op == 9 ? fp_canon<Q>(fp_sub_lazy<Q>(x, y)) : op == 10 ? fp_canon<Q>(fp_add_lazy<Q>(x, y)) : op == 11 ? fp_canon<Q>(fp_sqr_lazy<Q>(x)) : op == 12 ? fp_sqr_lazy<Q>(x) : op == 13 ? fp_canon<Q>(fp_mul_lazy_kara<Q>(x, y)) : op == 14 ? fp_mul_lazy_kara<Q>(x, y) : fp_neg<Q>(x);
        }
        st(out + 32 * i, r);
    }
    return 0;
}

// (a b + c d) R^-1 through the fused pair of products; inputs may be lazy (< 2p, or 2p itself); output canonicalised
int hc_mul2(int tag, const uint8_t* a, const uint8_t* b, const uint8_t* c, const uint8_t* d, uint8_t* out, uint64_t n, int raw) {
    for (uint64_t i = 0; i < n; ++i) {
        fp_t r = tag == 0 ? fp_mul2_lazy<S>(ld(a + 32 * i), ld(b + 32 * i), ld(c + 32 * i), ld(d + 32 * i))
                          : fp_mul2_lazy<Q>(ld(a + 32 * i), ld(b + 32 * i), ld(c + 32 * i), ld(d + 32 * i));
        if (!raw) r = tag == 0 ? fp_canon<S>(r) : fp_canon<Q>(r);
        st(out + 32 * i, r);
    }
    return 0;
}
// 2p - a
int hc_neg_lazy(int tag, const uint8_t* a, uint8_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) st(out + 32 * i, tag == 0 ? fp_neg_lazy<S>(ld(a + 32 * i)) : fp_neg_lazy<Q>(ld(a + 32 * i)));
    return 0;
}

// mode 0: madd(from_affine(a), b) ; 1: full add of two non-trivially scaled points ; 2: dbl(a) ; 3: a + b via madd into (a+b)-b
int hc_g1_op(int mode, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    affine_t A = lda(a), B = lda(b);
    xyzz_t r;
    if (mode == 0) {
        r = xyzz_from_affine(A);
        xyzz_madd(r, B);
    } else if (mode == 1) {
        // give both operands zz != 1: A' = 2A - A , B' = 2B - B
        xyzz_t a2 = xyzz_dbl(xyzz_from_affine(A)), b2 = xyzz_dbl(xyzz_from_affine(B));
        xyzz_madd(a2, affine_neg(A));
        xyzz_madd(b2, affine_neg(B));
        r = xyzz_add(a2, b2);
    } else if (mode == 2) {
        r = xyzz_dbl(xyzz_from_affine(A));
    } else if (mode == 3) {
        r = xyzz_from_affine(A);
        xyzz_madd(r, B);
        xyzz_madd(r, affine_neg(B));
        xyzz_madd(r, B);
    } else if (mode == 4) {
        // the kernels' hot path: lazily reduced accumulator ([0, 2p) coordinates), canonicalised once at the end.
        // A + B - B + B + A - A  exercises addition, and the doubling / cancellation checks on lazy values
        r = xyzz_from_affine(A);
        xyzz_madd_hot(r, B);
        xyzz_madd_hot(r, affine_neg(B));
        xyzz_madd_hot(r, B);
        xyzz_madd_hot(r, A);
        xyzz_madd_hot(r, affine_neg(A));
        xyzz_canon(r);
    } else {
        // mode 5: 2A via the doubling branch of the hot path, then + B, then (2A + B) - (2A + B) = infinity + B
        r = xyzz_from_affine(A);
        xyzz_madd_hot(r, A);
        xyzz_madd_hot(r, B);
        xyzz_t t = r;
        xyzz_canon(t);
        affine_t s = xyzz_to_affine(t);
        xyzz_madd_hot(r, affine_neg(s));
        xyzz_madd_hot(r, B);
        xyzz_canon(r);
    }
    sta(out, xyzz_to_affine(r));
    return 0;
}

// sum_i k_i P_i with signed c-bit windows through recode_signed + madd/dbl/add (exercises all group code)
int hc_msm_windowed(const uint8_t* bases, const uint8_t* scalars, uint64_t n, int c, uint8_t* out) {
    int W = num_windows(c);
    std::vector<std::vector<int32_t>> dig(n, std::vector<int32_t>(W));
    for (uint64_t i = 0; i < n; ++i) {
        fp_t k = fp_from_mont<S>(ld(scalars + 32 * i));
        recode_signed(k.l, c, W, dig[i].data());
    }
    xyzz_t acc = xyzz_inf();
    for (int w = W - 1; w >= 0; --w) {
        for (int b = 0; b < c; ++b) acc = xyzz_dbl(acc);
        xyzz_t wsum = xyzz_inf();
        for (uint64_t i = 0; i < n; ++i) {
            int32_t d = dig[i][w];
            if (d == 0) continue;
            affine_t P = lda(bases + 64 * i);
            xyzz_t t = xyzz_mul_u32(xyzz_from_affine(P), (uint32_t)(d < 0 ? -d : d));
            if (d < 0) t = xyzz_neg(t);
            wsum = xyzz_add(wsum, t);
        }
        acc = xyzz_add(acc, wsum);
    }
    sta(out, xyzz_to_affine(acc));
    return 0;
}

int hc_compress(const uint8_t* pts, uint64_t n, uint8_t* out32) {
    for (uint64_t i = 0; i < n; ++i) affine_serialize(lda(pts + 64 * i), out32 + 32 * i);
    return 0;
}
int hc_sha256(const uint8_t* msg, uint32_t len, uint8_t* out) {
    sha256_ctx c;
    sha256_init(c);
    sha256_update(c, msg, len);
    sha256_final(c, out);
    return 0;
}
int hc_xmd48(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len, uint32_t zpad, uint8_t* out48) {
    expand_message_xmd48(msg, len, dst, dst_len, zpad, out48);
    return 0;
}
int hc_hash_to_fr(const uint8_t* msg, uint32_t len, const uint8_t* dst, uint32_t dst_len, uint8_t* out) {
    st(out, hash_to_fr(msg, len, dst, dst_len));
    return 0;
}
int hc_fr_from_le32(const uint8_t* b, uint8_t* out) {
    st(out, fr_from_le32_mod_order(b));
    return 0;
}
// transcript walk used by the IPA opening: prefix, C, input point, output point -> w ; then (L,R) -> x
int hc_transcript_ipa(const uint8_t* prefix, uint32_t prefix_len, const char* dst, const uint8_t* C, const uint8_t* z, const uint8_t* y,
                      const uint8_t* L, const uint8_t* R, uint8_t* w_out, uint8_t* x_out) {
    transcript_t t;
    t.len = 0;
    t.mid_bytes = 0;
    tr_append_raw(t, prefix, prefix_len);
    t.dst_len = (uint32_t)strlen(dst);
    memcpy(t.dst, dst, t.dst_len);
    tr_append_point(t, lda(C), "C");
    tr_append_fr(t, ld(z), "input point");
    tr_append_fr(t, ld(y), "output point");
    st(w_out, tr_digest(t, "w"));
    tr_append_point(t, lda(L), "L");
    tr_append_point(t, lda(R), "R");
    st(x_out, tr_digest(t, "x"));
    return 0;
}
}
