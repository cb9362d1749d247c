// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the arkworks-0.4 arithmetic the reference (SleepingShell/verkle-kzg)
// relies on: BN254 Fr/Fq Montgomery fields (R = 2^256), G1 Jacobian group law,
// ark-serialize compressed encoding, SHA-256 + RFC 9380 expand_message_xmd as used by
// ark-ff 0.4 DefaultFieldHasher.  arkworks is an un-vendored dependency
// (vector-commit/Cargo.toml:12-26: ark-ff/ark-ec/ark-poly/ark-bn254 = "0.4",
// ark-serialize = "0.4.2", sha2 = "0.10.7"); its source is not on disk, and the reference
// holds no golden vectors, so:  **PARITY UNPINNED at the arkworks boundary.**
// What IS pinned (tests/test_oracle_*.py): SHA-256 FIPS vectors, RFC 9380 XMD vectors,
// BN254 2G/3G known coordinates, an independent pure-Python big-int model
// (tests/pyref.py), and the reference's own prove->verify properties.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this code.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <array>

namespace orc {

typedef unsigned __int128 u128;

// ----------------------------------------------------------------------------------
// 256-bit helpers (4 x u64, little-endian limbs)
// ----------------------------------------------------------------------------------
struct U256 {
    uint64_t l[4];
};

static inline int u256_cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; --i) {
        if (a.l[i] < b.l[i]) return -1;
        if (a.l[i] > b.l[i]) return 1;
    }
    return 0;
}
static inline uint64_t u256_add(U256& r, const U256& a, const U256& b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline uint64_t u256_sub(U256& r, const U256& a, const U256& b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static inline bool u256_is_zero(const U256& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline int u256_bit(const U256& a, int i) { return (int)((a.l[i >> 6] >> (i & 63)) & 1); }
static inline int u256_bitlen(const U256& a) {
    for (int i = 255; i >= 0; --i)
        if (u256_bit(a, i)) return i + 1;
    return 0;
}

// ----------------------------------------------------------------------------------
// Prime field in Montgomery form, R = 2^256 (bit-identical to arkworks' Fp256<MontBackend<_,4>>).
// Tag selects the modulus: 0 = Fr (scalar field), 1 = Fq (base field).
// ----------------------------------------------------------------------------------
struct FieldParams {
    U256 p;        // modulus
    U256 r1;       // R mod p   (Montgomery form of 1)
    U256 r2;       // R^2 mod p
    uint64_t inv;  // -p^{-1} mod 2^64
    U256 pm2;      // p - 2 (Fermat inversion exponent)
    U256 half;     // (p - 1) / 2
};

FieldParams make_params(const uint64_t p[4]);
const FieldParams& params(int tag);

template <int TAG>
struct Fp {
    U256 v;  // Montgomery representation

    static const FieldParams& P() { return params(TAG); }
    static Fp zero() {
        Fp r;
        memset(&r, 0, sizeof r);
        return r;
    }
    static Fp one() {
        Fp r;
        r.v = P().r1;
        return r;
    }
    bool is_zero() const { return u256_is_zero(v); }
    bool operator==(const Fp& o) const { return u256_cmp(v, o.v) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    Fp operator+(const Fp& o) const {
        Fp r;
        uint64_t c = u256_add(r.v, v, o.v);
        if (c || u256_cmp(r.v, P().p) >= 0) u256_sub(r.v, r.v, P().p);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        if (u256_sub(r.v, v, o.v)) u256_add(r.v, r.v, P().p);
        return r;
    }
    Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    Fp dbl() const { return *this + *this; }

    // CIOS Montgomery multiplication, 4 x 64-bit limbs.
    Fp operator*(const Fp& o) const {
        const FieldParams& pr = P();
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)v.l[j] * o.v.l[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * pr.inv;
            c = (u128)m * pr.p.l[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)m * pr.p.l[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp r;
        memcpy(r.v.l, t, 32);
        if (t[4] || u256_cmp(r.v, pr.p) >= 0) u256_sub(r.v, r.v, pr.p);
        return r;
    }
    Fp sqr() const { return *this * *this; }
    Fp& operator+=(const Fp& o) { return *this = *this + o; }
    Fp& operator-=(const Fp& o) { return *this = *this - o; }
    Fp& operator*=(const Fp& o) { return *this = *this * o; }

    Fp pow(const U256& e) const {
        Fp acc = one();
        int n = u256_bitlen(e);
        for (int i = n - 1; i >= 0; --i) {
            acc = acc.sqr();
            if (u256_bit(e, i)) acc = acc * *this;
        }
        return acc;
    }
    Fp pow_u64(uint64_t e) const {
        U256 x = {{e, 0, 0, 0}};
        return pow(x);
    }
    // ark-ff Field::inverse(): None for zero.  Here zero -> zero (callers check).
    Fp inverse() const { return pow(P().pm2); }

    // canonical (non-Montgomery) integer
    U256 to_canonical() const {
        Fp one_raw;
        one_raw.v = U256{{1, 0, 0, 0}};
        return (*this * one_raw).v;
    }
    static Fp from_canonical(const U256& x) {  // x must be < p
        Fp a, r2;
        a.v = x;
        r2.v = P().r2;
        return a * r2;
    }
    static Fp from_u64(uint64_t x) { return from_canonical(U256{{x, 0, 0, 0}}); }

    // ark-ff PrimeField::from_{le,be}_bytes_mod_order — plain big-integer reduction.
    static Fp from_be_bytes_mod_order(const uint8_t* b, size_t n) {
        Fp acc = zero();
        Fp k256 = from_u64(256);
        for (size_t i = 0; i < n; ++i) acc = acc * k256 + from_u64(b[i]);
        return acc;
    }
    static Fp from_le_bytes_mod_order(const uint8_t* b, size_t n) {
        Fp acc = zero();
        Fp k256 = from_u64(256);
        for (size_t i = n; i-- > 0;) acc = acc * k256 + from_u64(b[i]);
        return acc;
    }
    // ark-serialize: canonical integer, 32 bytes little-endian.
    void serialize(uint8_t out[32]) const {
        U256 c = to_canonical();
        memcpy(out, c.l, 32);
    }
    // Ord on ark-ff Fp compares canonical integers (precompute.rs:74, lagrange_basis.rs:64-66).
    int cmp(const Fp& o) const { return u256_cmp(to_canonical(), o.to_canonical()); }
};

typedef Fp<0> Fr;
typedef Fp<1> Fq;

// ark_ff::batch_inversion: zeros are skipped (left as zero).
template <class F>
void batch_inversion(std::vector<F>& v) {
    std::vector<F> prod;
    prod.reserve(v.size());
    F acc = F::one();
    for (auto& x : v) {
        if (!x.is_zero()) {
            acc = acc * x;
            prod.push_back(acc);
        }
    }
    acc = acc.inverse();
    size_t k = prod.size();
    for (size_t i = v.size(); i-- > 0;) {
        if (v[i].is_zero()) continue;
        --k;
        F prev = (k == 0) ? F::one() : prod[k - 1];
        F inv = acc * prev;
        acc = acc * v[i];
        v[i] = inv;
    }
}

// ----------------------------------------------------------------------------------
// G1: y^2 = x^3 + 3 over Fq, Jacobian coordinates as ark-ec 0.4 short_weierstrass::Projective.
// ----------------------------------------------------------------------------------
struct G1Affine {
    Fq x, y;
    bool infinity;
};

struct G1 {
    Fq X, Y, Z;  // Z == 0 <=> identity

    static G1 identity() {
        G1 r;
        r.X = Fq::one();
        r.Y = Fq::one();
        r.Z = Fq::zero();
        return r;
    }
    static G1 generator() {
        G1 r;
        r.X = Fq::from_u64(1);
        r.Y = Fq::from_u64(2);
        r.Z = Fq::one();
        return r;
    }
    static G1 from_affine(const G1Affine& a) {
        if (a.infinity) return identity();
        G1 r;
        r.X = a.x;
        r.Y = a.y;
        r.Z = Fq::one();
        return r;
    }
    bool is_zero() const { return Z.is_zero(); }

    G1 dbl() const {
        if (is_zero()) return *this;
        // dbl-2009-l (a = 0)
        Fq A = X.sqr(), B = Y.sqr(), C = B.sqr();
        Fq D = ((X + B).sqr() - A - C).dbl();
        Fq E = A + A + A;
        Fq F = E.sqr();
        G1 r;
        r.X = F - D.dbl();
        r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
        r.Z = (Y * Z).dbl();
        return r;
    }
    G1 operator+(const G1& o) const {
        if (is_zero()) return o;
        if (o.is_zero()) return *this;
        // add-2007-bl
        Fq Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        Fq U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        Fq S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return identity();
        }
        Fq H = U2 - U1;
        Fq I = H.dbl().sqr();
        Fq J = H * I;
        Fq rr = (S2 - S1).dbl();
        Fq V = U1 * I;
        G1 r;
        r.X = rr.sqr() - J - V.dbl();
        r.Y = rr * (V - r.X) - (S1 * J).dbl();
        r.Z = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
        return r;
    }
    G1 neg() const {
        G1 r = *this;
        r.Y = Y.neg();
        return r;
    }
    G1 operator-(const G1& o) const { return *this + o.neg(); }
    G1& operator+=(const G1& o) { return *this = *this + o; }

    // Group * ScalarField (utils.rs:17): MSB-first double-and-add over the canonical scalar,
    // the algorithm ark-ec's mul_bigint uses (one double per bit, one add per set bit).
    G1 mul(const Fr& k) const {
        U256 e = k.to_canonical();
        G1 acc = identity();
        int n = u256_bitlen(e);
        for (int i = n - 1; i >= 0; --i) {
            acc = acc.dbl();
            if (u256_bit(e, i)) acc = acc + *this;
        }
        return acc;
    }
    G1Affine to_affine() const {
        G1Affine a;
        if (is_zero()) {
            a.x = Fq::zero();
            a.y = Fq::zero();
            a.infinity = true;
            return a;
        }
        Fq zi = Z.inverse();
        Fq zi2 = zi.sqr();
        a.x = X * zi2;
        a.y = Y * zi2 * zi;
        a.infinity = false;
        return a;
    }
    bool operator==(const G1& o) const {
        if (is_zero() || o.is_zero()) return is_zero() && o.is_zero();
        Fq Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        if (X * Z2Z2 != o.X * Z1Z1) return false;
        return Y * o.Z * Z2Z2 == o.Y * Z * Z1Z1;
    }
    bool operator!=(const G1& o) const { return !(*this == o); }
};

bool g1_on_curve(const G1Affine& a);

// ark-ec 0.4 SWCurveConfig::serialize_with_mode(Compress::Yes):
//   x as 32-byte LE canonical integer; byte 31 |= 0x80 if y > -y (SWFlags::YIsNegative),
//   |= 0x40 for the point at infinity (x = 0).          [UNVERIFIED-HERE: arkworks not on disk]
void g1_serialize_compressed(const G1& p, uint8_t out[32]);
void g1_affine_serialize_compressed(const G1Affine& a, uint8_t out[32]);
// ark-ec 0.4 `Affine::from_random_bytes` (short Weierstrass) on 32 bytes, as EthereumHashToCurve calls it
// (ipa/ipa_point_generator.rs:103-104): Fq::from_random_bytes_with_flags::<SWFlags> takes the two top bits of byte 31 as
// flags, masks them off and rejects x >= p; flags 11 -> None; infinity flag -> identity iff x == 0, else None; otherwise
// y = sqrt(x^3 + 3) (None for a non-residue) and the flag picks between y and -y by the SAME convention as the compressed
// encoding above (bit 7 set <=> the y that serialises with bit 7), so from_random_bytes(serialize(P)) == P.  No subgroup
// check (cofactor 1).  Returns false for None.                       [UNVERIFIED-HERE: same constant as the encoder]
bool g1_affine_from_random_bytes(const uint8_t bytes[32], G1Affine& out);

// ----------------------------------------------------------------------------------
// SHA-256 (FIPS 180-4) and the ark-ff 0.4 DefaultFieldHasher<Sha256, 128>
// ----------------------------------------------------------------------------------
void sha256(const uint8_t* msg, size_t len, uint8_t out[32]);

// The ONE place the arkworks-0.4 Z_pad wrinkle lives: ark-ff 0.4 builds ExpanderXmd with
// block_size = len_per_base_elem (48 for BN254 Fr at 128-bit security) instead of SHA-256's
// 64-byte input block.  RFC 9380 vectors are checked with z_pad_len = 64.
static const size_t ARK04_Z_PAD_LEN = 48;
static const size_t ARK04_LEN_PER_ELEM = 48;  // ceil((254 + 128) / 8)

std::vector<uint8_t> expand_message_xmd(const uint8_t* msg, size_t msg_len, const uint8_t* dst,
                                        size_t dst_len, size_t len_in_bytes, size_t z_pad_len);
// hash_to_field(msg, 1)[0] for Fr
Fr hash_to_fr(const uint8_t* msg, size_t msg_len, const std::string& dst);

// transcript.rs:34-62  TranscriptHasher
struct Transcript {
    std::vector<uint8_t> state;
    std::string dst;
    explicit Transcript(const std::string& label) : dst(label) {}
    void append_bytes(const uint8_t* b, size_t n, const std::string& label) {
        state.insert(state.end(), label.begin(), label.end());
        state.insert(state.end(), b, b + n);
    }
    void append(const G1& p, const std::string& label) {
        uint8_t b[32];
        g1_serialize_compressed(p, b);
        append_bytes(b, 32, label);
    }
    void append(const Fr& f, const std::string& label) {
        uint8_t b[32];
        f.serialize(b);
        append_bytes(b, 32, label);
    }
    void append_usize(uint64_t z, const std::string& label) {  // usize -> u64 LE
        uint8_t b[8];
        memcpy(b, &z, 8);
        append_bytes(b, 8, label);
    }
    Fr digest(const std::string& label, bool clear) {
        state.insert(state.end(), label.begin(), label.end());
        Fr res = hash_to_fr(state.data(), state.size(), dst);
        if (clear) {
            uint8_t b[32];
            res.serialize(b);
            state.assign(b, b + 32);
            state.insert(state.end(), label.begin(), label.end());
        }
        return res;
    }
};

// ark-poly Radix2EvaluationDomain::new(n): size = next pow2, group_gen = 5^((r-1)/size).
Fr domain_group_gen(uint64_t size_pow2);
static inline uint64_t next_pow2(uint64_t n) {
    uint64_t s = 1;
    while (s < n) s <<= 1;
    return s;
}

}  // namespace orc
