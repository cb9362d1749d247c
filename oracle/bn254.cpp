// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254.hpp header).
#include "bn254.hpp"

namespace orc {

// BN254 moduli (ark-bn254 0.4: fr.rs / fq.rs)
static const uint64_t FR_MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                                   0x30644e72e131a029ULL};
static const uint64_t FQ_MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL,
                                   0x30644e72e131a029ULL};

FieldParams make_params(const uint64_t p[4]) {
    FieldParams fp;
    memcpy(fp.p.l, p, 32);
    // inv = -p^{-1} mod 2^64 by Newton iteration
    uint64_t x = 1;
    for (int i = 0; i < 7; ++i) x *= 2 - p[0] * x;
    fp.inv = (uint64_t)(0 - x);
    // R mod p and R^2 mod p by repeated modular doubling of 1
    U256 acc = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; ++i) {
        U256 d;
        uint64_t c = u256_add(d, acc, acc);
        if (c || u256_cmp(d, fp.p) >= 0) u256_sub(d, d, fp.p);
        acc = d;
        if (i == 255) fp.r1 = acc;
    }
    fp.r2 = acc;
    U256 two = {{2, 0, 0, 0}}, one = {{1, 0, 0, 0}};
    u256_sub(fp.pm2, fp.p, two);
    U256 pm1;
    u256_sub(pm1, fp.p, one);
    for (int i = 0; i < 4; ++i) fp.half.l[i] = (pm1.l[i] >> 1) | (i < 3 ? (pm1.l[i + 1] << 63) : 0);
    return fp;
}

const FieldParams& params(int tag) {
    static const FieldParams fr = make_params(FR_MOD);
    static const FieldParams fq = make_params(FQ_MOD);
    return tag == 0 ? fr : fq;
}

bool g1_on_curve(const G1Affine& a) {
    if (a.infinity) return true;
    return a.y.sqr() == a.x.sqr() * a.x + Fq::from_u64(3);
}

void g1_affine_serialize_compressed(const G1Affine& a, uint8_t out[32]) {
    if (a.infinity) {
        memset(out, 0, 32);
        out[31] |= 0x40;  // SWFlags::PointAtInfinity
        return;
    }
    U256 xc = a.x.to_canonical();
    memcpy(out, xc.l, 32);
    U256 yc = a.y.to_canonical();
    // SWFlags::from_y_coordinate: y <= -y -> positive ; else YIsNegative (bit 7)
    if (u256_cmp(yc, Fq::P().half) > 0) out[31] |= 0x80;
}

bool g1_affine_from_random_bytes(const uint8_t bytes[32], G1Affine& out) {
    uint8_t b[32];
    memcpy(b, bytes, 32);
    const uint8_t flags = b[31] & 0xC0;
    b[31] &= 0x3F;
    U256 xc;
    memcpy(xc.l, b, 32);
    if (u256_cmp(xc, Fq::P().p) >= 0) return false;  // deserialize_compressed of the masked bytes fails
    if (flags == 0xC0) return false;                 // SWFlags::from_u8: both bits set is invalid
    if (flags == 0x40) {                             // infinity flag
        if (!u256_is_zero(xc)) return false;
        out.infinity = true;
        out.x = Fq::zero();
        out.y = Fq::zero();
        return true;
    }
    Fq x = Fq::from_canonical(xc);
    Fq rhs = x.sqr() * x + Fq::from_u64(3);
    // p = 3 mod 4: the candidate root is rhs^((p+1)/4)
    U256 e, one = {{1, 0, 0, 0}};
    u256_add(e, Fq::P().p, one);
    for (int i = 0; i < 4; ++i) e.l[i] = (e.l[i] >> 2) | (i < 3 ? (e.l[i + 1] << 62) : 0);
    Fq y = rhs.pow(e);
    if (!(y.sqr() == rhs)) return false;
    const bool y_is_larger = u256_cmp(y.to_canonical(), Fq::P().half) > 0;
    const bool want_larger = flags == 0x80;  // the encoder sets bit 7 for the larger root
    out.infinity = false;
    out.x = x;
    out.y = (y_is_larger == want_larger) ? y : y.neg();
    return true;
}

void g1_serialize_compressed(const G1& p, uint8_t out[32]) { g1_affine_serialize_compressed(p.to_affine(), out); }

// ---------------------------------------------------------------- SHA-256
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(uint32_t h[8], const uint8_t* blk) {
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t full = len / 64;
    for (size_t i = 0; i < full; ++i) sha256_block(h, msg + 64 * i);
    uint8_t tail[128];
    size_t rem = len - full * 64;
    memset(tail, 0, sizeof tail);
    if (rem) memcpy(tail, msg + full * 64, rem);
    tail[rem] = 0x80;
    size_t tl = (rem + 9 <= 64) ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(h, tail);
    if (tl == 128) sha256_block(h, tail + 64);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8);
        out[4 * i + 3] = (uint8_t)h[i];
    }
}

// RFC 9380 section 5.3.1 expand_message_xmd with SHA-256; z_pad_len is a parameter because of
// the ark-ff 0.4 wrinkle (see header).  DST longer than 255 bytes is not needed on this path.
std::vector<uint8_t> expand_message_xmd(const uint8_t* msg, size_t msg_len, const uint8_t* dst, size_t dst_len,
                                        size_t n, size_t z_pad_len) {
    const size_t b_len = 32;
    size_t ell = (n + b_len - 1) / b_len;
    std::vector<uint8_t> dst_prime(dst, dst + dst_len);
    dst_prime.push_back((uint8_t)dst_len);

    std::vector<uint8_t> m0(z_pad_len, 0);
    m0.insert(m0.end(), msg, msg + msg_len);
    m0.push_back((uint8_t)(n >> 8));
    m0.push_back((uint8_t)n);
    m0.push_back(0);
    m0.insert(m0.end(), dst_prime.begin(), dst_prime.end());
    uint8_t b0[32], bi[32];
    sha256(m0.data(), m0.size(), b0);

    std::vector<uint8_t> m1(b0, b0 + 32);
    m1.push_back(1);
    m1.insert(m1.end(), dst_prime.begin(), dst_prime.end());
    sha256(m1.data(), m1.size(), bi);

    std::vector<uint8_t> out(bi, bi + 32);
    for (size_t i = 2; i <= ell; ++i) {
        std::vector<uint8_t> mi(32);
        for (int k = 0; k < 32; ++k) mi[k] = b0[k] ^ bi[k];
        mi.push_back((uint8_t)i);
        mi.insert(mi.end(), dst_prime.begin(), dst_prime.end());
        sha256(mi.data(), mi.size(), bi);
        out.insert(out.end(), bi, bi + 32);
    }
    out.resize(n);
    return out;
}

Fr hash_to_fr(const uint8_t* msg, size_t msg_len, const std::string& dst) {
    std::vector<uint8_t> u = expand_message_xmd(msg, msg_len, (const uint8_t*)dst.data(), dst.size(), ARK04_LEN_PER_ELEM,
                                                ARK04_Z_PAD_LEN);
    return Fr::from_be_bytes_mod_order(u.data(), u.size());
}

Fr domain_group_gen(uint64_t size_pow2) {
    // 5^((r-1)/size)
    U256 e, one = {{1, 0, 0, 0}};
    u256_sub(e, Fr::P().p, one);
    int lg = 0;
    while ((1ULL << lg) < size_pow2) ++lg;
    for (int s = 0; s < lg; ++s)
        for (int i = 0; i < 4; ++i) e.l[i] = (e.l[i] >> 1) | (i < 3 ? (e.l[i + 1] << 63) : 0);
    return Fr::from_u64(5).pow(e);
}

}  // namespace orc
