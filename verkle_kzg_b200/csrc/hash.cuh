// Fiat-Shamir transcript of the reference (vector-commit/src/transcript.rs:34-62) for the device:
// SHA-256, RFC 9380 expand_message_xmd as ark-ff 0.4's DefaultFieldHasher<Sha256,128> drives it
// (Z_pad = 48 zero bytes, 48 output bytes, big-endian reduction mod r), and the byte-string state
// machine (append = label || compressed bytes; digest = label, hash, state <- ser(res) || label).
// Host/device so that tests/host can check it byte for byte against hashlib on the CPU.
#pragma once
#include "curve.cuh"

namespace vk {

struct sha256_ctx {
    uint32_t h[8];
    uint8_t buf[64];
    uint32_t buflen;
    uint64_t total;
};

VK_HD uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

// out of line: the fully unrolled 64 rounds are ~4 KB of code and the transcript kernels call it from many sites
__host__ __device__ __noinline__ inline void sha256_compress_words(uint32_t h[8], const uint32_t blk[16]) {
    const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = blk[i];
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K[i] + wi;
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

__host__ __device__ inline void sha256_compress(uint32_t h[8], const uint8_t* blk) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    sha256_compress_words(h, w);
}

VK_HD void sha256_init(sha256_ctx& c) {
    c.h[0] = 0x6a09e667; c.h[1] = 0xbb67ae85; c.h[2] = 0x3c6ef372; c.h[3] = 0xa54ff53a;
    c.h[4] = 0x510e527f; c.h[5] = 0x9b05688c; c.h[6] = 0x1f83d9ab; c.h[7] = 0x5be0cd19;
    c.buflen = 0;
    c.total = 0;
}
__host__ __device__ inline void sha256_update(sha256_ctx& c, const uint8_t* p, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) {
        c.buf[c.buflen++] = p[i];
        if (c.buflen == 64) {
            sha256_compress(c.h, c.buf);
            c.buflen = 0;
        }
    }
    c.total += n;
}
__host__ __device__ inline void sha256_update_zeros(sha256_ctx& c, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) {
        c.buf[c.buflen++] = 0;
        if (c.buflen == 64) {
            sha256_compress(c.h, c.buf);
            c.buflen = 0;
        }
    }
    c.total += n;
}
__host__ __device__ inline void sha256_final(sha256_ctx& c, uint8_t out[32]) {
    uint64_t bits = c.total * 8;
    uint8_t pad = 0x80;
    sha256_update(c, &pad, 1);
    while (c.buflen != 56) sha256_update_zeros(c, 1);
    uint8_t lenb[8];
    for (int i = 0; i < 8; ++i) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_update(c, lenb, 8);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(c.h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(c.h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(c.h[i] >> 8);
        out[4 * i + 3] = (uint8_t)c.h[i];
    }
}

// ark-ff 0.4 DefaultFieldHasher<Sha256,128>::hash_to_field(msg, 1)[0] for BN254 Fr.
// The arkworks-0.4 wrinkle (Z_pad = len_per_base_elem = 48 rather than the 64-byte SHA block) is this
// single constant; RFC 9380 vectors are checked in tests/host with z_pad_len = 64.
static const uint32_t ARK04_Z_PAD_LEN = 48;

// big-endian word i of a 64-byte block held as bytes in 16 aligned words
VK_HD uint32_t be_word(const uint32_t* blk_words, int i) {
#ifdef __CUDA_ARCH__
    return __byte_perm(blk_words[i], 0, 0x0123);
#else
    const uint8_t* p = reinterpret_cast<const uint8_t*>(blk_words + i);
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
#endif
}
// copy the part of the segment [seg0, seg1) (bytes src[0 .. seg1 - seg0)) that falls into the block starting at `base`
VK_HD void blk_copy(uint8_t* blk, uint32_t base, uint32_t seg0, uint32_t seg1, const uint8_t* src) {
    uint32_t lo = seg0 > base ? seg0 : base, hi = seg1 < base + 64 ? seg1 : base + 64;
    for (uint32_t q = lo; q < hi; ++q) blk[q - base] = src[q - seg0];
}
VK_HD void blk_put(uint8_t* blk, uint32_t base, uint32_t pos, uint32_t byte) {
    if (pos >= base && pos < base + 64) blk[pos - base] = (uint8_t)byte;
}

// expand_message_xmd (RFC 9380 5.3.1) for 48 output bytes = b1 || b2[0..16), on 32-bit words.  On the device a
// byte-at-a-time SHA update (one buffer-full test and one length update per byte) is a chain of branches and
// local-memory round trips, ~90 cycles per byte for a lone thread — more than the compressions themselves.  Here every
// 64-byte block is zeroed, filled by a few straight segment copies, and read back as words; b1 / b2 take b0 (and
// b0 ^ b1) from the state words.
// `mid` / `mid_bytes` (optional): the SHA-256 state after the first mid_bytes (a multiple of 64, >= z_pad_len) bytes of
// Z_pad || msg were absorbed elsewhere (a long in-flight transcript is pre-hashed on the host); `msg` then holds the
// message bytes from that position on.
__host__ __device__ inline void xmd48_words(const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len,
                                            uint32_t z_pad_len, uint32_t b1[8], uint32_t b2[8], const uint32_t* mid = nullptr,
                                            uint32_t mid_bytes = 0) {
    const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    // b0 = H(Z_pad || msg || I2OSP(48, 2) || I2OSP(0, 1) || DST || I2OSP(len(DST), 1))
    const uint32_t m0 = mid_bytes ? mid_bytes : z_pad_len, m1 = m0 + msg_len, d0 = m1 + 3, d1 = d0 + dst_len, total = d1 + 1;
    const uint32_t nblk = (total + 9 + 63) / 64;
    uint32_t b0[8], w[16], bw[16];
    uint8_t* bb = reinterpret_cast<uint8_t*>(bw);
    for (int i = 0; i < 8; ++i) b0[i] = mid_bytes ? mid[i] : IV[i];
    for (uint32_t blk = mid_bytes / 64; blk < nblk; ++blk) {
        const uint32_t base = blk * 64;
#pragma unroll
        for (int i = 0; i < 16; ++i) bw[i] = 0;
        blk_copy(bb, base, m0, m1, msg);
        blk_put(bb, base, m1 + 1, 48);
        blk_copy(bb, base, d0, d1, dst);
        blk_put(bb, base, d1, dst_len);
        blk_put(bb, base, total, 0x80);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = be_word(bw, i);
        if (blk == nblk - 1) w[15] = total * 8;  // (message lengths here are far below 2^29 bytes)
        sha256_compress_words(b0, w);
    }
    // b_i = H(strxor(b0, b_(i-1)) || I2OSP(i, 1) || DST || I2OSP(len(DST), 1)),  b_0' = 0  (one block for DSTs <= 21 bytes)
    const uint32_t tl = 32 + 1 + dst_len + 1, tblk = (tl + 9 + 63) / 64;
    for (int round = 1; round <= 2; ++round) {
        uint32_t* out = round == 1 ? b1 : b2;
        uint32_t h[8];
        for (int i = 0; i < 8; ++i) h[i] = IV[i];
        for (uint32_t blk = 0; blk < tblk; ++blk) {
            const uint32_t base = blk * 64;
#pragma unroll
            for (int i = 0; i < 16; ++i) bw[i] = 0;
            blk_put(bb, base, 32, (uint32_t)round);
            blk_copy(bb, base, 33, 33 + dst_len, dst);
            blk_put(bb, base, 33 + dst_len, dst_len);
            blk_put(bb, base, tl, 0x80);
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = be_word(bw, i);
            if (blk == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = round == 1 ? b0[i] : (b0[i] ^ b1[i]);
            }
            if (blk == tblk - 1) w[15] = tl * 8;
            sha256_compress_words(h, w);
        }
        for (int i = 0; i < 8; ++i) out[i] = h[i];
    }
}

__host__ __device__ inline void expand_message_xmd48(const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len,
                                                     uint32_t z_pad_len, uint8_t out[48]) {
    uint32_t b1[8], b2[8];
    xmd48_words(msg, msg_len, dst, dst_len, z_pad_len, b1, b2);
    for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 4; ++k) out[4 * i + k] = (uint8_t)(b1[i] >> (24 - 8 * k));
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 4; ++k) out[32 + 4 * i + k] = (uint8_t)(b2[i] >> (24 - 8 * k));
}

// Fr (Montgomery) from 48 big-endian bytes, reduced mod r:  v = hi * 2^256 + lo
//   mont(v) = lo * R^2 / R + hi * R^3 / R
__host__ __device__ inline fp_t fr_from_be48(const uint8_t u[48]) {
    const uint32_t R3[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
    fp_t hi, lo, r2, r3;
    for (int i = 0; i < 8; ++i) {
        // limb i (little-endian) of lo = bytes u[16 + 28 - 4i .. +4) big-endian
        const uint8_t* p = u + 16 + 28 - 4 * i;
        lo.l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
        r2.l[i] = S::r2(i);
        r3.l[i] = R3[i];
        hi.l[i] = 0;
    }
    for (int i = 0; i < 4; ++i) {
        const uint8_t* p = u + 12 - 4 * i;
        hi.l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    // fp_mul needs only its FIRST operand < r; the second may be any 256-bit value
    return fp_add<S>(fp_mul_ni<S>(r2, lo), fp_mul_ni<S>(r3, hi));
}

// the same from the big-endian words of b1 || b2[0..4): hi = b1[0..4), lo = b1[4..8) || b2[0..4)
__host__ __device__ inline fp_t fr_from_be48_words(const uint32_t b1[8], const uint32_t b2[8]) {
    const uint32_t R3[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu};
    fp_t hi, lo, r2, r3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r2.l[i] = S::r2(i);
        r3.l[i] = R3[i];
        hi.l[i] = i < 4 ? b1[3 - i] : 0;
        lo.l[i] = i < 4 ? b2[3 - i] : b1[11 - i];
    }
    return fp_add<S>(fp_mul_ni<S>(r2, lo), fp_mul_ni<S>(r3, hi));
}

__host__ __device__ inline fp_t hash_to_fr(const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len,
                                           const uint32_t* mid = nullptr, uint32_t mid_bytes = 0) {
    uint32_t b1[8], b2[8];
    xmd48_words(msg, msg_len, dst, dst_len, ARK04_Z_PAD_LEN, b1, b2, mid, mid_bytes);
    return fr_from_be48_words(b1, b2);
}

// Fr -> 32 little-endian canonical bytes (ark-serialize)
__host__ __device__ inline void fr_serialize(const fp_t& a_mont, uint8_t out[32]) {
    fp_t c = fp_from_mont<S>(a_mont);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)c.l[i];
        out[4 * i + 1] = (uint8_t)(c.l[i] >> 8);
        out[4 * i + 2] = (uint8_t)(c.l[i] >> 16);
        out[4 * i + 3] = (uint8_t)(c.l[i] >> 24);
    }
}
__host__ __device__ inline void affine_serialize(const affine_t& p, uint8_t out[32]) {
    uint32_t w[8];
    affine_compress(p, w);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)w[i];
        out[4 * i + 1] = (uint8_t)(w[i] >> 8);
        out[4 * i + 2] = (uint8_t)(w[i] >> 16);
        out[4 * i + 3] = (uint8_t)(w[i] >> 24);
    }
}
// Fr from 32 little-endian bytes reduced mod r (from_le_bytes_mod_order on a 32-byte string): the value
// is < 2^256 = R, so one Montgomery multiplication by R^2 both reduces and converts.
__host__ __device__ inline fp_t fr_from_le32_mod_order(const uint8_t b[32]) {
    fp_t v, r2;
    for (int i = 0; i < 8; ++i) {
        v.l[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
        r2.l[i] = S::r2(i);
    }
    return fp_mul_ni<S>(r2, v);  // first operand < r, second any 256-bit value
}

// Transcript state: at most TR_MAX bytes held as bytes (an in-flight prefix of up to TR_PREFIX_INLINE bytes, the IPA
// opening adds <= 121 before the first clearing digest, every later round holds 100).  A LONGER in-flight prefix — the
// reference's transcript has no size limit — is pre-hashed on the host: `mid` is the SHA-256 state after the first
// `mid_bytes` bytes of Z_pad || state, and `state` holds what follows (make_prefix in ipa.cu).
static const uint32_t TR_MAX = 320;
static const uint32_t TR_PREFIX_INLINE = 160;
static const uint32_t TR_DST_MAX = 16;

struct transcript_t {
    uint8_t state[TR_MAX];
    uint32_t len;
    uint8_t dst[TR_DST_MAX];
    uint32_t dst_len;
    uint32_t mid[8];
    uint32_t mid_bytes;  // 0: nothing pre-hashed
};

__host__ __device__ inline void tr_append_raw(transcript_t& t, const uint8_t* p, uint32_t n) {
    const uint32_t at = t.len, room = TR_MAX - at;
    if (n > room) n = room;
    for (uint32_t i = 0; i < n; ++i) t.state[at + i] = p[i];
    t.len = at + n;
}
__host__ __device__ inline void tr_append_label(transcript_t& t, const char* label) {
    for (const char* p = label; *p; ++p)
        if (t.len < TR_MAX) t.state[t.len++] = (uint8_t)*p;
}
__host__ __device__ inline void tr_append_point(transcript_t& t, const affine_t& p, const char* label) {
    uint8_t b[32];
    tr_append_label(t, label);
    affine_serialize(p, b);
    tr_append_raw(t, b, 32);
}
__host__ __device__ inline void tr_append_fr(transcript_t& t, const fp_t& x, const char* label) {
    uint8_t b[32];
    tr_append_label(t, label);
    fr_serialize(x, b);
    tr_append_raw(t, b, 32);
}
// digest(label, clear = true)
__host__ __device__ inline fp_t tr_digest(transcript_t& t, const char* label) {
    tr_append_label(t, label);
    fp_t res = hash_to_fr(t.state, t.len, t.dst, t.dst_len, t.mid, t.mid_bytes);
    uint8_t b[32];
    fr_serialize(res, b);
    t.len = 0;
    t.mid_bytes = 0;
    tr_append_raw(t, b, 32);
    tr_append_label(t, label);
    return res;
}

}  // namespace vk
