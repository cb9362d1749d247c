// Host side of the OUTER multiproof transcript (multiproof.rs:108-115): SHA-256 over ~75 bytes per query of
// caller-supplied host data, plus the canonical serialisation of the (C, y) it absorbs.  A strictly serial hash
// chain is the one piece of this path a GPU thread is bad at (~2.5 us per 64-byte block), so it runs here while
// the rows upload: SHA-NI when the CPU has it, portable C otherwise; Montgomery -> canonical with 64-bit limbs.
#include <cstdint>
#include <cstring>
#include <cstddef>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void compress_portable(uint32_t h[8], const uint8_t* p, size_t blocks) {
    while (blocks--) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        p += 64;
    }
}

#if defined(__x86_64__)
__attribute__((target("sha,sse4.1,ssse3"))) void compress_shani(uint32_t state[8], const uint8_t* data, size_t blocks) {
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TMP = _mm_loadu_si128((const __m128i*)&state[0]);
    __m128i STATE1 = _mm_loadu_si128((const __m128i*)&state[4]);
    TMP = _mm_shuffle_epi32(TMP, 0xB1);
    STATE1 = _mm_shuffle_epi32(STATE1, 0x1B);
    __m128i STATE0 = _mm_alignr_epi8(TMP, STATE1, 8);
    STATE1 = _mm_blend_epi16(STATE1, TMP, 0xF0);
    while (blocks--) {
        __m128i ABEF = STATE0, CDGH = STATE1, MSG, M[4];
        for (int i = 0; i < 4; ++i) M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(data + 16 * i)), MASK);
        for (int r = 0; r < 16; ++r) {
            __m128i cur = M[r & 3];
            MSG = _mm_add_epi32(cur, _mm_loadu_si128((const __m128i*)&K256[4 * r]));
            STATE1 = _mm_sha256rnds2_epu32(STATE1, STATE0, MSG);
            MSG = _mm_shuffle_epi32(MSG, 0x0E);
            STATE0 = _mm_sha256rnds2_epu32(STATE0, STATE1, MSG);
            if (r < 12) {  // schedule the message words of round r + 4
                __m128i t = _mm_sha256msg1_epu32(M[r & 3], M[(r + 1) & 3]);
                t = _mm_add_epi32(t, _mm_alignr_epi8(M[(r + 3) & 3], M[(r + 2) & 3], 4));
                M[r & 3] = _mm_sha256msg2_epu32(t, M[(r + 3) & 3]);
            }
        }
        STATE0 = _mm_add_epi32(STATE0, ABEF);
        STATE1 = _mm_add_epi32(STATE1, CDGH);
        data += 64;
    }
    TMP = _mm_shuffle_epi32(STATE0, 0x1B);
    STATE1 = _mm_shuffle_epi32(STATE1, 0xB1);
    STATE0 = _mm_blend_epi16(TMP, STATE1, 0xF0);
    STATE1 = _mm_alignr_epi8(STATE1, TMP, 8);
    _mm_storeu_si128((__m128i*)&state[0], STATE0);
    _mm_storeu_si128((__m128i*)&state[4], STATE1);
}
#endif

typedef void (*compress_fn)(uint32_t*, const uint8_t*, size_t);
compress_fn pick() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3")) return compress_shani;
#endif
    return compress_portable;
}

// 4 x 64-bit limb Montgomery reduction of `a` (multiplication by 1): a * R^-1 mod p, fully reduced
void redc(const uint64_t a[4], const uint64_t p[4], uint64_t inv, uint64_t out[4]) {
    uint64_t t[5] = {a[0], a[1], a[2], a[3], 0};
    for (int i = 0; i < 4; ++i) {
        uint64_t m = t[0] * inv;
        unsigned __int128 c = (unsigned __int128)m * p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (unsigned __int128)m * p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = (uint64_t)(c >> 64);
    }
    // t < 2p: conditional subtraction
    uint64_t d[4];
    unsigned __int128 b = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 x = (unsigned __int128)t[i] - p[i] - (uint64_t)b;
        d[i] = (uint64_t)x;
        b = (x >> 64) & 1;
    }
    bool ge = t[4] != 0 || b == 0;
    for (int i = 0; i < 4; ++i) out[i] = ge ? d[i] : t[i];
}

const uint64_t FR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
const uint64_t FQ_P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
const uint64_t FR_INV = 0xc2e1f593efffffffULL;  // -r^-1 mod 2^64
const uint64_t FQ_INV = 0x87d20782e4866389ULL;  // -p^-1 mod 2^64
const uint64_t FQ_HALF[4] = {0x9e10460b6c3e7ea3ULL, 0xcbc0b548b438e546ULL, 0xdc2822db40c0ac2eULL, 0x183227397098d014ULL};  // (p-1)/2

// 4 x 64-bit limb Montgomery product a * b * R^-1 mod r (CIOS); a may be any 256-bit value, b < r
void montmul_fr(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (unsigned __int128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV;
        c = (unsigned __int128)m * FR_P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (unsigned __int128)m * FR_P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    uint64_t d[4];
    unsigned __int128 bw = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 x = (unsigned __int128)t[i] - FR_P[i] - (uint64_t)bw;
        d[i] = (uint64_t)x;
        bw = (x >> 64) & 1;
    }
    bool ge = t[4] != 0 || bw == 0;
    for (int i = 0; i < 4; ++i) out[i] = ge ? d[i] : t[i];
}

const uint64_t FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};  // R^2 mod r

}  // namespace

extern "C" {

// F::from_le_bytes_mod_order(bytes) in Montgomery form (the literals of verkle nodes: 16-byte value halves, stems)
void vkh_fr_from_le_bytes(const uint8_t* b, size_t len, uint64_t out[4]) {
    // Horner over 31-byte digits (each < 2^248 < r), most significant first: acc = acc * 2^248 + digit
    uint64_t acc[4] = {0, 0, 0, 0};
    uint64_t radix[4] = {0, 0, 0, 1ULL << 56}, radix_m[4];
    montmul_fr(radix, FR_R2, radix_m);  // 2^248 in Montgomery form
    size_t digits = (len + 30) / 31;
    for (size_t d = digits; d-- > 0;) {
        size_t lo = d * 31, n = len - lo < 31 ? len - lo : 31;
        uint8_t tmp[32] = {0};
        memcpy(tmp, b + lo, n);
        uint64_t v[4], vm[4];
        memcpy(v, tmp, 32);
        montmul_fr(v, FR_R2, vm);
        if (d + 1 < digits) montmul_fr(acc, radix_m, acc);
        // acc += vm (mod r)
        unsigned __int128 c = 0;
        uint64_t s[4];
        for (int i = 0; i < 4; ++i) {
            c += (unsigned __int128)acc[i] + vm[i];
            s[i] = (uint64_t)c;
            c >>= 64;
        }
        uint64_t dd[4];
        unsigned __int128 bw = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 x = (unsigned __int128)s[i] - FR_P[i] - (uint64_t)bw;
            dd[i] = (uint64_t)x;
            bw = (x >> 64) & 1;
        }
        bool ge = c != 0 || bw == 0;
        for (int i = 0; i < 4; ++i) acc[i] = ge ? dd[i] : s[i];
    }
    memcpy(out, acc, 32);
}

// one-shot SHA-256 of a contiguous buffer
void vkh_sha256(const uint8_t* data, size_t len, uint8_t out[32]) {
    static compress_fn fn = pick();
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t full = len / 64;
    fn(h, data, full);
    uint8_t tail[128];
    size_t rem = len - full * 64;
    memset(tail, 0, sizeof(tail));
    memcpy(tail, data + full * 64, rem);
    tail[rem] = 0x80;
    size_t tb = rem + 9 <= 64 ? 1 : 2;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) tail[tb * 64 - 1 - i] = (uint8_t)(bits >> (8 * i));
    fn(h, tail, tb);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8);
        out[4 * i + 3] = (uint8_t)h[i];
    }
}

// ark-serialize of an Fr: canonical little-endian 32 bytes
void vkh_serialize_fr(const uint64_t mont[4], uint8_t out[32]) {
    uint64_t c[4];
    redc(mont, FR_P, FR_INV, c);
    memcpy(out, c, 32);
}

// ark-serialize compressed G1: canonical x, bit 255 = y > (p-1)/2, bit 254 (x = 0) = infinity
void vkh_serialize_g1(const uint64_t xy_mont[8], uint8_t out[32]) {
    bool inf = true;
    for (int i = 0; i < 8; ++i) inf = inf && xy_mont[i] == 0;
    if (inf) {
        memset(out, 0, 32);
        out[31] = 0x40;
        return;
    }
    uint64_t x[4], y[4];
    redc(xy_mont, FQ_P, FQ_INV, x);
    redc(xy_mont + 4, FQ_P, FQ_INV, y);
    bool gt = false;
    for (int i = 3; i >= 0; --i) {
        if (y[i] != FQ_HALF[i]) {
            gt = y[i] > FQ_HALF[i];
            break;
        }
    }
    memcpy(out, x, 32);
    if (gt) out[31] |= 0x80;
}

int vkh_has_shani(void) { return pick() != compress_portable; }

}  // extern "C"
