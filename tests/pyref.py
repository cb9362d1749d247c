"""Independent pure-Python big-int model of the reference's hot path (second implementation).

TEST INFRASTRUCTURE.  Used only to cross-check the C++ oracle (oracle/) at small sizes and to
generate tests/golden/*.json.  Written directly from the reference's Rust sources
(/root/reference/vector-commit/src/...; each function cites file:line) with Python ints and
hashlib, sharing no code with oracle/ or the CUDA product.

arkworks-0.4 conventions restated here (UNVERIFIED against a real arkworks build: none is
available offline): see SURVEY.md section 8(c) items 1-6.
"""
import hashlib

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
MONT_R = 1 << 256
G1_GEN = (1, 2)
Z_PAD_LEN_ARK04 = 48      # ark-ff 0.4: ExpanderXmd.block_size = len_per_base_elem
LEN_PER_ELEM = 48         # ceil((254 + 128) / 8)


# ------------------------------------------------------------------ fields
def finv(x, m=R_MOD):
    return pow(x, m - 2, m)


# ------------------------------------------------------------------ G1 (affine tuples; None = identity)
def g_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * finv(2 * y1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * finv(x2 - x1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return (x3, (lam * (x1 - x3) - y1) % P_MOD)


def g_neg(a):
    return None if a is None else (a[0], (-a[1]) % P_MOD)


def g_sub(a, b):
    return g_add(a, g_neg(b))


def g_mul(a, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = g_add(acc, a)
        a = g_add(a, a)
        k >>= 1
    return acc


def g_on_curve(a):
    return a is None or (a[1] * a[1] - a[0] ** 3 - 3) % P_MOD == 0


def inner_product_g(bases, scalars):
    """utils.rs:16-19 (zip truncates)."""
    acc = None
    for p, s in zip(bases, scalars):
        acc = g_add(acc, g_mul(p, s))
    return acc


def inner_product_f(a, b):
    return sum(x * y for x, y in zip(a, b)) % R_MOD


# ------------------------------------------------------------------ serialisation (ark-serialize 0.4)
def ser_fr(x):
    return int(x % R_MOD).to_bytes(32, "little")


def ser_g1(a):
    """Compressed SW point: x LE, byte31 |= 0x80 if y > -y, 0x40 for infinity."""
    if a is None:
        b = bytearray(32)
        b[31] |= 0x40
        return bytes(b)
    b = bytearray(a[0].to_bytes(32, "little"))
    if a[1] > (P_MOD - a[1]) % P_MOD:
        b[31] |= 0x80
    return bytes(b)


def from_random_bytes(b):
    """ark-ec 0.4 Affine::from_random_bytes on 32 bytes (same flag convention as ser_g1): point, None-identity
    is returned as the string "inf", rejection as False."""
    flags = b[31] & 0xC0
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x3F]), "little")
    if x >= P_MOD or flags == 0xC0:
        return False
    if flags == 0x40:
        return "inf" if x == 0 else False
    rhs = (x * x * x + 3) % P_MOD
    y = pow(rhs, (P_MOD + 1) // 4, P_MOD)
    if y * y % P_MOD != rhs:
        return False
    larger = max(y, P_MOD - y)
    return (x, larger if flags == 0x80 else P_MOD - larger)


def ipa_crs_gen(seed, num):
    """ipa_point_generator.rs:51-70 with EthereumHashToCurve (:97-109)"""
    import hashlib
    res, i = [], 0
    while len(res) < num:
        pt = from_random_bytes(hashlib.sha256(bytes(seed) + i.to_bytes(8, "little")).digest())
        if pt is not False:
            res.append(None if pt == "inf" else pt)
        i += 1
    return res, i


def to_data_item(a):
    """lib.rs:56-67."""
    if a is None:
        return 0
    return int.from_bytes(ser_g1(a), "little") % R_MOD


# ------------------------------------------------------------------ hash to field (ark-ff 0.4 DefaultFieldHasher<Sha256,128>)
def expand_message_xmd(msg, dst, n, z_pad_len=64):
    ell = (n + 31) // 32
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(z_pad_len) + msg + n.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:n]


def hash_to_fr(msg, dst):
    u = expand_message_xmd(msg, dst, LEN_PER_ELEM, Z_PAD_LEN_ARK04)
    return int.from_bytes(u, "big") % R_MOD


class Transcript:
    """transcript.rs:34-62."""

    def __init__(self, label, state=b""):
        self.dst = label.encode()
        self.state = bytes(state)

    def append_g(self, p, label):
        self.state += label.encode() + ser_g1(p)

    def append_f(self, x, label):
        self.state += label.encode() + ser_fr(x)

    def append_usize(self, z, label):
        self.state += label.encode() + int(z).to_bytes(8, "little")

    def digest(self, label, clear=True):
        self.state += label.encode()
        res = hash_to_fr(self.state, self.dst)
        if clear:
            self.state = ser_fr(res) + label.encode()
        return res


# ------------------------------------------------------------------ domain / precompute.rs
def next_pow2(n):
    s = 1
    while s < n:
        s <<= 1
    return s


def group_gen(n):
    return pow(5, (R_MOD - 1) // next_pow2(n), R_MOD)


def vanishing_evaluations(N):
    """precompute.rs:46-58: A'(w^i) = N * w^-i."""
    w = group_gen(N)
    ev = [N * finv(pow(w, i, R_MOD)) % R_MOD for i in range(N)]
    return ev, [finv(e) for e in ev]


def barycentric(N, point):
    """precompute.rs:72-90."""
    res = [0] * N
    if point < N:
        res[point] = 1
        return res
    w = group_gen(N)
    t = (pow(point, N, R_MOD) - 1) * finv(N) % R_MOD
    for i in range(N):
        pw = pow(w, i, R_MOD)
        res[i] = t * pw % R_MOD * finv((point - pw) % R_MOD) % R_MOD
    return res


# ------------------------------------------------------------------ lagrange_basis.rs
def evaluate(N, data, domain_n, point):
    """lagrange_basis.rs:63-72 (max = len(data))."""
    mx = len(data) - 1
    if point <= mx:
        return data[point]
    if point <= next_pow2(domain_n):
        return 0
    return inner_product_f(data, barycentric(N, point))


def divide_by_vanishing(N, data, domain_n, index):
    """lagrange_basis.rs:91-119."""
    n = next_pow2(domain_n)
    w = group_gen(n)
    ev, evinv = vanishing_evaluations(N)
    q = [0] * n
    index_f = pow(w, index, R_MOD)
    val = data[index] if index < len(data) else 0
    for i in range(n):
        if i == index:
            continue
        i_f = pow(w, i, R_MOD)
        i_eval = data[i] if i < len(data) else 0
        sub = (i_eval - val) % R_MOD
        q[i] = sub * finv((i_f - index_f) % R_MOD) % R_MOD
        q[index] = (q[index] + sub * ev[index] % R_MOD * evinv[i] % R_MOD * finv((index_f - i_f) % R_MOD)) % R_MOD
    return q


def divide_by_vanishing_outside(N, data, domain_n, point):
    """lagrange_basis.rs:121-142."""
    n = next_pow2(domain_n)
    w = group_gen(n)
    val = evaluate(N, data, domain_n, point)
    return [((data[i] if i < len(data) else 0) - val) * finv((pow(w, i, R_MOD) - point) % R_MOD) % R_MOD for i in range(n)]


# ------------------------------------------------------------------ ipa/mod.rs
def low_level_ipa(gens, q, a, b, commitment, input_point, tr=None):
    """ipa/mod.rs:268-319. returns (L, R, tip, y)."""
    ev = inner_product_f(a, b)
    gens = list(gens[: len(a)])
    data, other = list(a), list(b)
    tr = tr or Transcript("ipa")
    tr.append_g(commitment, "C")
    tr.append_f(input_point, "input point")
    tr.append_f(ev, "output point")
    L, R = [], []
    ra = tr.digest("w")
    q = g_mul(q, ra)
    while len(data) > 1:
        h = len(data) // 2
        dl, dr = data[:h], data[h:]
        gl, gr = gens[: len(gens) // 2], gens[len(gens) // 2:]
        bl, br = other[: len(other) // 2], other[len(other) // 2:]
        yl = g_add(inner_product_g(gr, dl), g_mul(q, inner_product_f(dl, br)))
        yr = g_add(inner_product_g(gl, dr), g_mul(q, inner_product_f(dr, bl)))
        L.append(yl)
        R.append(yr)
        tr.append_g(yl, "L")
        tr.append_g(yr, "R")
        ra = tr.digest("x")
        data = [(x + ra * y) % R_MOD for x, y in zip(dl, dr)]
        gens = [g_add(x, g_mul(y, ra)) for x, y in zip(gr, gl)]
        other = [(x + ra * y) % R_MOD for x, y in zip(br, bl)]
    return L, R, data[0], ev


def low_level_verify_ipa(gens, q, b, commitment, input_point, proof, tr=None):
    """ipa/mod.rs:321-360."""
    L, R, tip, y = proof
    c = commitment
    tr = tr or Transcript("ipa")
    tr.append_g(commitment, "C")
    tr.append_f(input_point, "input point")
    tr.append_f(y, "output point")
    ra = tr.digest("w")
    coeffs = [1]
    q = g_mul(q, ra)
    c = g_add(c, g_mul(q, y))
    for l, r in zip(L, R):
        tr.append_g(l, "L")
        tr.append_g(r, "R")
        ra = tr.digest("x")
        c = g_add(g_add(l, g_mul(c, ra)), g_mul(r, ra * ra % R_MOD))
        coeffs = [v for x in coeffs for v in (x * ra % R_MOD, x)]
    cp = inner_product_g(gens, coeffs)
    cb = inner_product_f(b, coeffs)
    return c == g_add(g_mul(cp, tip), g_mul(q, tip * cb % R_MOD))


def ipa_prove_point(g, q, N, a, commitment, point, tr=None):
    return low_level_ipa(g, q, a, barycentric(N, point), commitment, point, tr)


def ipa_verify_point(g, q, N, commitment, point, proof, tr=None):
    return low_level_verify_ipa(g, q, barycentric(N, point), commitment, point, proof, tr)


# ------------------------------------------------------------------ kzg/mod.rs
def kzg_setup(n, tau):
    """kzg_point_generator.rs:32-43 + kzg/mod.rs:115-124 (n a power of two here)."""
    w = group_gen(n)
    out = []
    for j in range(n):
        wj = finv(pow(w, j, R_MOD))
        lj = sum(pow(tau, i, R_MOD) * pow(wj, i, R_MOD) for i in range(n)) % R_MOD * finv(n) % R_MOD
        out.append(g_mul(G1_GEN, lj))
    return out


def kzg_prove_point(lag, data, point):
    """kzg/mod.rs:136-154. returns (proof, y)."""
    N = len(lag)
    y = evaluate(N, data, N, point)
    if point <= N:
        q = divide_by_vanishing(N, data, N, point)
    else:
        q = divide_by_vanishing_outside(N, data, N, point)
    return inner_product_g(lag, q), y


def kzg_verify_tau(lag, tau, commitment, point, proof):
    """kzg/mod.rs:165-189 on the scalar side (tau known)."""
    pi, y = proof
    N = len(lag)
    p = pow(group_gen(N), point, R_MOD) if point < N else point
    return g_mul(pi, (tau - p) % R_MOD) == g_sub(commitment, g_mul(G1_GEN, y))


# ------------------------------------------------------------------ multiproof.rs
def multiproof_prove(scheme, bases, N, queries):
    """multiproof.rs:99-176.  queries: list of (data, commit, z, y).  scheme 'ipa' (bases N+1) or 'kzg'."""
    tr = Transcript("multiproof")
    for data, c, z, y in queries:
        tr.append_g(c, "C")
        tr.append_usize(z, "z")
        tr.append_f(y, "y")
    r = tr.digest("r")
    by_point = {}
    rp = 1
    for data, c, z, y in queries:
        by_point.setdefault(z, []).append([x * rp % R_MOD for x in data])
        rp = rp * r % R_MOD
    g = [0] * N
    for z, rows in by_point.items():
        total = [sum(col) % R_MOD for col in zip(*rows)]
        quo = divide_by_vanishing(N, total, N, z)
        g = [(x + y) % R_MOD for x, y in zip(g, quo)]
    d = inner_product_g(bases, g)
    tr.append_g(d, "D")
    t = tr.digest("t")
    inv = [finv((t - i) % R_MOD) for i in range(N)]
    h = [0] * N
    for z, rows in by_point.items():
        for row in rows:
            h = [(x + y * inv[z]) % R_MOD for x, y in zip(h, row)]
    e = inner_product_g(bases, h)
    tr.append_g(e, "E")
    hmg = [(x - y) % R_MOD for x, y in zip(h, g)]
    if scheme == "ipa":
        proof = ipa_prove_point(bases[:N], bases[N], N, hmg, g_sub(e, d), t, tr)
    else:
        proof = kzg_prove_point(bases, hmg, t)
    return proof, d


def multiproof_verify(scheme, bases, N, vqueries, proof, d, tau=None):
    """multiproof.rs:178-215.  vqueries: list of (commit, z, y)."""
    tr = Transcript("multiproof")
    for c, z, y in vqueries:
        tr.append_g(c, "C")
        tr.append_usize(z, "z")
        tr.append_f(y, "y")
    r = tr.digest("r")
    tr.append_g(d, "D")
    t = tr.digest("t")
    inv = [finv((t - i) % R_MOD) for i in range(N)]
    coeffs = {}
    rp = 1
    for c, z, y in vqueries:
        coeffs[c] = (coeffs.get(c, 0) + rp * inv[z]) % R_MOD
        rp = rp * r % R_MOD
    e = None
    for c, k in coeffs.items():
        e = g_add(e, g_mul(c, k))
    tr.append_g(e, "E")
    if scheme == "ipa":
        return ipa_verify_point(bases[:N], bases[N], N, g_sub(e, d), t, proof, tr)
    return kzg_verify_tau(bases, tau, g_sub(e, d), t, proof)
