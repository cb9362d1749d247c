"""Randomised parity sweep (test infrastructure, not collected by pytest): libvkzg against the oracle on random shapes
and seeds.      python tests/fuzz_parity.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import orc  # noqa: E402
from verkle_kzg_b200 import Engine  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    eng = Engine(0)
    rng = np.random.default_rng(int(time.time()))
    t0 = time.time()
    it = 0
    while time.time() - t0 < budget:
        it += 1
        N = int(rng.choice([2, 4, 8, 16, 32, 64, 256]))
        c = int(rng.choice([4, 7, 8, 11, 13, 16]))
        k0, k1 = orc.rand_fr(rng, 2)
        bases = orc.points_walk(k0, k1, N + 1)
        key = eng.load_key(bases[:N], q=bases[N], window_bits=c)
        B = int(rng.integers(1, 40))
        w = int(rng.integers(1, N + 1))
        # scalars: mix of uniform, small, sparse and edge values
        s = orc.rand_fr_buf(rng, B * w).reshape(B, w, 32)
        mask = rng.random((B, w)) < 0.3
        s[mask] = 0
        small = rng.random((B, w)) < 0.2
        s[small] = orc.fr_to_buf([int(rng.integers(0, 1 << 20))])[0]
        got = eng.commit_batch(key, s)
        assert (got == orc.commit_batch(bases[:N], s)).all(), ("commit", N, c, B, w)
        # IPA
        Bp = int(rng.integers(1, 6))
        a = orc.rand_fr_buf(rng, Bp * N).reshape(Bp, N, 32)
        C = eng.commit_batch(key, a)
        zs = [int(rng.integers(0, 3 * N)) if rng.random() < 0.7 else orc.rand_fr(rng, 1)[0] for _ in range(Bp)]
        zs = [z for z in zs]
        zb = orc.fr_to_buf(zs)
        L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C)
        for i in range(Bp):
            eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
            assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all(), ("ipa", N, c, zs[i])
        assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
        key.free()
        # MSM
        n = int(rng.integers(1, 3000))
        cm = int(rng.choice([6, 9, 12, 16]))
        mb = orc.points_walk(k1, k0, n)
        mk = eng.load_key(mb, kind=2, window_bits=cm)
        ms = orc.rand_fr_buf(rng, n)
        ms[rng.random(n) < 0.2] = 0
        assert (eng.msm(mk, ms) == orc.msm(mb, ms, mode="pippenger")).all(), ("msm", n, cm)
        mk.free()
    print(f"fuzz ok: {it} iterations in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
