"""Host-side mirror of the reference's operator interface for the commitment path — the traits
VectorCommitment (vector-commit/src/lib.rs:70-174) and VectorCommitmentMultiproof
(vector-commit/src/multiproof.rs:90-216) — over libvkzg.  Same method names, argument meaning and error
behaviour; every method is a thin marshalling layer over one C-ABI call (no arithmetic happens here).

Types (numpy uint8 in the ABI layouts):
    Data        LagrangeBasis: evaluations [len, 32]                      (lagrange_basis.rs:14-21)
    Commitment  [64]  affine G1, all-zero = identity                      (ark Projective on the Rust side)
    IPAProof    dict(l [log2 N, 64], r [log2 N, 64], tip [32], y [32])    (ipa/mod.rs:79-84)
    KZGProof    dict(proof [64], y [32])                                  (kzg/mod.rs:81-84)
    Multiproof  dict(proof, d [64])                                       (multiproof.rs:55-58)
The reference's methods are associated functions taking `&UniversalParams`; here the params object also
carries the engine (GPU context) that holds the key's tables.
"""
import numpy as np

from ._lib import KEY_WINDOW, VkzgError

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_MONT_R = 1 << 256


def fr_from_int(v):
    """F::from(v) in the ABI layout"""
    return np.frombuffer(((int(v) % R_MOD) * _MONT_R % R_MOD).to_bytes(32, "little"), dtype=np.uint8).copy()


class OutOfDomain(Exception):
    """KZGError::OutOfDomain / IPAError::OutOfDomain (kzg/mod.rs:88, ipa/mod.rs:88)"""


class OutOfCRS(Exception):
    """KZGError::OutOfCRS / IPAError::OutOfCRS"""


class OutOfBounds(Exception):
    """PointGeneratorError::OutOfBounds (lib.rs:176-182)"""


class InvalidPoint(Exception):
    """PointGeneratorError::InvalidPoint"""


class IPAPointGenerator:
    """IPAPointGenerator<G, EthereumHashToCurve> (ipa/ipa_point_generator.rs:14-86): same constructor arguments,
    defaults (max 256, seed "eth_verkle_oct_2021"), bounds checks and error cases; the hashing and the square roots
    run on the GPU (vkzg_ipa_crs_generate)."""

    def __init__(self, engine, max=256, seed=b"eth_verkle_oct_2021"):
        self.engine = engine
        self.max = max
        self.seed = bytes(seed)

    def set_max(self, max):
        self.max = max

    def gen(self, num):
        if num > self.max:
            raise OutOfBounds()
        if num == 0:
            return np.zeros((0, 64), dtype=np.uint8)
        return self.engine.ipa_crs_generate(self.seed, num)[0]

    def gen_at(self, index):
        if index > self.max:  # the reference's own (inclusive) bound, ipa_point_generator.rs:73
            raise OutOfBounds()
        p = self.engine.ipa_crs_generate_at(self.seed, index)
        if p is None:
            raise InvalidPoint()
        return p

    def secret(self):
        return self.seed


class LagrangeBasis:
    """VCData impl of the reference (lagrange_basis.rs:151-178)"""

    def __init__(self, evaluations, domain_n=0):
        self.evaluations = np.ascontiguousarray(evaluations, dtype=np.uint8).reshape(-1, 32)
        self.domain_n = domain_n  # 0: D::new(len)

    @classmethod
    def from_vec(cls, data):
        return cls(data)

    @classmethod
    def from_vec_and_domain(cls, data, domain_n):
        """lagrange_basis.rs:24-31 with D::new(domain_n)"""
        return cls(data, domain_n)

    def max(self):
        return len(self.evaluations) - 1

    def __len__(self):
        return len(self.evaluations)


class UniversalParams:
    """KZGKey (kzg/mod.rs:27-57) / IPAUniversalParams (ipa/mod.rs:22-52): bases resident on the GPU"""

    def __init__(self, engine, bases, q=None, window_bits=0):
        self.engine = engine
        self.key = engine.load_key(bases, q=q, kind=KEY_WINDOW, window_bits=window_bits)
        self.size = self.key.n

    def max_size(self):
        return self.size

    def free(self):
        self.key.free()


class _Scheme:
    @staticmethod
    def setup(engine, bases, q=None, window_bits=0):
        """VectorCommitment::setup with the points already generated (the PointGenerators are setup code
        outside the hot path)"""
        return UniversalParams(engine, bases, q, window_bits)

    @staticmethod
    def commit(key, data):
        """commit (kzg/mod.rs:126-134, ipa/mod.rs:130-135): inner_product zips bases and data (quirk Q1)"""
        ev = data.evaluations
        if len(ev) > key.size:
            ev = ev[: key.size]
        if len(ev) == 0:
            return np.zeros(64, dtype=np.uint8)
        return key.engine.commit_batch(key.key, ev.reshape(1, -1, 32))[0]

    @classmethod
    def commit_batch(cls, key, datas):
        """B commitments of equal width in one launch: datas [B, w, 32]"""
        return key.engine.commit_batch(key.key, np.ascontiguousarray(datas, dtype=np.uint8))

    @classmethod
    def prove(cls, key, commitment, index, data):
        """VectorCommitment::prove (lib.rs:111-124): point = F::from(index)"""
        return cls.prove_point(key, commitment, fr_from_int(index), data)

    @classmethod
    def verify(cls, key, commitment, index, proof):
        """VectorCommitment::verify (lib.rs:136-149)"""
        return cls.verify_point(key, commitment, fr_from_int(index), proof)


class KZGRandomPointGenerator:
    """KZGRandomPointGenerator (kzg/kzg_point_generator.rs:10-52): the secret tau (default 100) and G * tau^i on the GPU"""

    def __init__(self, engine, secret=100):
        self.engine = engine
        self._secret = int(secret) % R_MOD

    def secret(self):
        return self._secret

    def _gen_key(self):
        g = np.zeros((1, 64), dtype=np.uint8)
        g[0, :32] = np.frombuffer((_MONT_R % _P_MOD).to_bytes(32, "little"), dtype=np.uint8)          # x = 1
        g[0, 32:] = np.frombuffer((2 * _MONT_R % _P_MOD).to_bytes(32, "little"), dtype=np.uint8)      # y = 2
        return self.engine.load_key(g, kind=KEY_WINDOW, window_bits=16)

    def gen(self, num):
        k = self._gen_key()
        try:
            return self.engine.kzg_powers(k, fr_from_int(self._secret), num)
        finally:
            k.free()


class IPA(_Scheme):
    """IPA<N, G, H, D> (ipa/mod.rs:98-181)"""

    @staticmethod
    def setup_from_generator(engine, max_items, gen, window_bits=0):
        """VectorCommitment::setup(max_items, gen) (ipa/mod.rs:121-128): gens = gen.gen(max_items + 1), g = the first
        max_items of them, q the next one.  `gen` is an IPAPointGenerator (its bound raises OutOfBounds like the reference)."""
        gens = gen.gen(max_items + 1)
        return UniversalParams(engine, gens[:max_items], gens[max_items], window_bits)

    @staticmethod
    def prove_point(key, commitment, point, data, transcript=None):
        """transcript: None or (state bytes, dst label) of an in-flight TranscriptHasher (lib.rs:127-133)"""
        if len(data) != key.size:
            raise VkzgError(-4, "IPA::prove_point: data width must equal the key width")
        prefix, dst = transcript if transcript else (b"", "ipa")
        L, R, tip, y = key.engine.ipa_prove_batch(key.key, data.evaluations.reshape(1, -1, 32), point.reshape(1, 32),
                                                  commitment.reshape(1, 64), prefix=prefix, dst=dst)
        return dict(l=L[0], r=R[0], tip=tip[0], y=y[0])

    @staticmethod
    def prove_point_batch(key, commitments, points, datas):
        L, R, tip, y = key.engine.ipa_prove_batch(key.key, datas, points, commitments)
        return dict(l=L, r=R, tip=tip, y=y)

    @staticmethod
    def verify_point(key, commitment, point, proof, transcript=None):
        prefix, dst = transcript if transcript else (b"", "ipa")
        ok = key.engine.ipa_verify_batch(key.key, point.reshape(1, 32), commitment.reshape(1, 64), proof["l"][None], proof["r"][None],
                                         proof["tip"][None], proof["y"][None], prefix=prefix, dst=dst)
        return bool(ok[0])

    @staticmethod
    def prove_commitment(key, commitment, data):
        """IPA::prove_commitment (ipa/mod.rs:199-235)"""
        L, R, tip = key.engine.ipa_prove_commitment_batch(key.key, data.evaluations.reshape(1, -1, 32), commitment.reshape(1, 64))
        return dict(l=L[0], r=R[0], tip=tip[0])

    @staticmethod
    def verify_commitment_proof(key, commitment, proof):
        """IPA::verify_commitment_proof (ipa/mod.rs:238-265)"""
        ok = key.engine.ipa_verify_commitment_batch(key.key, commitment.reshape(1, 64), proof["l"][None], proof["r"][None], proof["tip"][None])
        return bool(ok[0])

    @staticmethod
    def prove_multiproof(key, queries):
        """VectorCommitmentMultiproof::prove_multiproof (multiproof.rs:99-176).  queries: list of
        (data: LagrangeBasis, commit [64], z: int, y [32])"""
        return _prove_multiproof(key, queries, "ipa")

    @staticmethod
    def verify_multiproof(key, queries, proof):
        """verify_multiproof (multiproof.rs:178-215).  queries: list of (commit [64], z: int, y [32])"""
        C = np.stack([q[0] for q in queries])
        z = np.array([q[1] for q in queries], dtype=np.uint64)
        y = np.stack([q[2] for q in queries])
        p = proof["proof"]
        return key.engine.multiproof_verify_ipa(key.key, C, z, y, dict(D=proof["d"], L=p["l"], R=p["r"], tip=p["tip"], y=p["y"]))


class KZG(_Scheme):
    """KZG<E, H, D> (kzg/mod.rs:96-198)"""

    @staticmethod
    def setup_from_generator(engine, max_items, gen, window_bits=0):
        """VectorCommitment::setup(max_items, gen) (kzg/mod.rs:115-124): the Lagrange-form SRS over the radix-2 domain of size
        next_pow2(max_items).  The reference runs `domain.ifft` over gen.gen(max_items); with the generator's secret at hand
        (it is read there too, for the G2 element) the same canonical points come from closed-form scalars
        (vkzg_kzg_setup_from_secret).  The G2 element of KZGKey stays on the host side of the shim (pairings, K4)."""
        k = gen._gen_key()
        try:
            lagrange = engine.kzg_setup_from_secret(k, fr_from_int(gen.secret()), max_items)
        finally:
            k.free()
        return UniversalParams(engine, lagrange, None, window_bits)

    @staticmethod
    def prove_point(key, commitment, point, data, transcript=None):
        """kzg/mod.rs:136-154 (commitment and transcript are unused by the reference too)"""
        try:
            proof, y = key.engine.kzg_open_batch(key.key, data.evaluations.reshape(1, -1, 32), point.reshape(1, 32),
                                                 domain_n=data.domain_n)
        except VkzgError as e:
            if e.status == -3:
                raise OutOfDomain("point == key size: the reference indexes out of bounds here (quirk Q2)") from e
            raise
        return dict(proof=proof[0], y=y[0])

    @staticmethod
    def prove_point_batch(key, points, datas, domain_n=0):
        proof, y = key.engine.kzg_open_batch(key.key, datas, points, domain_n=domain_n)
        return dict(proof=proof, y=y)

    @staticmethod
    def prove_all_points(key, data):
        """kzg/mod.rs:200-235: the opening at every point of the data's domain, entry i equal to prove_point(key, _, i, data)"""
        proof, y = key.engine.kzg_prove_all_batch(key.key, data.evaluations.reshape(1, -1, 32), domain_n=data.domain_n)
        return [dict(proof=p, y=v) for p, v in zip(proof[0], y[0])]

    @staticmethod
    def verify_point(key, commitment, point, proof, transcript=None):
        raise NotImplementedError(
            "KZG::verify_point is two pairings (kzg/mod.rs:165-189); it stays on the host arkworks side of the "
            "shim (SURVEY.md section 8 row K4) and is not part of libvkzg")

    @staticmethod
    def prove_multiproof(key, queries):
        return _prove_multiproof(key, queries, "kzg")


def _prove_multiproof(key, queries, scheme):
    f = np.stack([q[0].evaluations for q in queries])
    C = np.stack([q[1] for q in queries])
    z = np.array([q[2] for q in queries], dtype=np.uint64)
    y = np.stack([q[3] for q in queries])
    if f.shape[1] != key.size:
        raise VkzgError(-4, "prove_multiproof: data width must equal the key width")
    out = key.engine.multiproof_prove(key.key, scheme, f, C, z, y)
    if scheme == "ipa":
        return dict(proof=dict(l=out["L"], r=out["R"], tip=out["tip"], y=out["y"]), d=out["D"])
    return dict(proof=dict(proof=out["L"][0], y=out["y"]), d=out["D"])
