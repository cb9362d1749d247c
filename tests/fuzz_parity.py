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
    tk0, tk1 = orc.rand_fr(rng, 2)
    tree_bases = orc.points_walk(tk0, tk1, 256)
    tree_key = eng.load_key(tree_bases, window_bits=8)
    t0 = time.time()
    it = 0
    while time.time() - t0 < budget:
        it += 1
        N = int(rng.choice([2, 4, 8, 16, 32, 64, 256]))
        c = int(rng.choice([4, 5, 7, 8, 11, 13, 15, 16, 17] if N <= 64 else [4, 7, 8, 11, 13, 16]))
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
        # multiproof over m queries of this key (IPA scheme; multiproof.rs:99-176), random group skew
        if N >= 4:
            m = int(rng.integers(1, 50))
            f = orc.rand_fr_buf(rng, m * N).reshape(m, N, 32)
            Cm = eng.commit_batch(key, f)
            z = rng.integers(0, N if rng.random() < 0.5 else 2, m).astype(np.uint64)
            ym = np.stack([f[i, int(z[i])] for i in range(m)])
            got = eng.multiproof_prove(key, "ipa", f, Cm, z, ym)
            exp = orc.multiproof_prove("ipa", bases, N, f, Cm, z, ym)
            assert all((got[k] == exp[k]).all() for k in ("D", "L", "R", "tip", "y")), ("multiproof", N, c, m)
            assert eng.multiproof_verify_ipa(key, Cm, z, ym, got)
        key.free()
        # KZG open over a Lagrange SRS (kzg/mod.rs:136-154): in-domain indices and outside points
        srs = orc.kzg_setup(N, 100)
        kk = eng.load_key(srs, window_bits=c)
        Bk = int(rng.integers(1, 8))
        fk = orc.rand_fr_buf(rng, Bk * N).reshape(Bk, N, 32)
        zk = [int(rng.integers(0, N)) if rng.random() < 0.5 else int(rng.integers(N + 1, 1 << 40)) for _ in range(Bk)]
        zkb = orc.fr_to_buf(zk)
        pf, yk = eng.kzg_open_batch(kk, fk, zkb)
        for i in range(Bk):
            epf, ey, ok = orc.kzg_prove(srs, fk[i], zkb[i])
            assert ok and (pf[i] == epf).all() and (yk[i] == ey).all(), ("kzg", N, c, zk[i])
        if it % 4 == 1:   # all-points prover (kzg/mod.rs:200-235 by its contract): entry i == prove_point at i
            pa, ya = eng.kzg_prove_all_batch(kk, fk[:2])
            for b_ in range(min(2, Bk)):
                for i in {0, N - 1, int(rng.integers(0, N))}:
                    epf, ey, ok = orc.kzg_prove(srs, fk[b_], orc.fr_to_buf([i])[0])
                    assert ok and (pa[b_, i] == epf).all() and (ya[b_, i] == ey).all(), ("kzg-all", N, c, i)
        kk.free()
        # native host tree against the oracle's replay of the same insertion sequence, random flatten mode
        if tree_key is not None and it % 4 == 0:
            from verkle_kzg_b200.tree import NativeVerkleTree, VerkleTree
            kl = int(rng.choice([3, 4, 8, 32, 40]))
            width = int(rng.choice([1, 2, 3, 32, 256]))
            nk = int(rng.integers(1, 400))
            keys = rng.integers(0, int(rng.choice([3, 256])), (nk, kl), dtype=np.uint8)
            _, first = np.unique(keys[:, : kl - 1], axis=0, return_index=True)
            keys = keys[np.sort(first)]
            vals = rng.integers(0, 256, (len(keys), 32), dtype=np.uint8)
            ref = VerkleTree(kl, 256)
            keep = 0
            for k_, v_ in zip(keys, vals):
                try:
                    ref.insert_single(k_, v_)
                    keep += 1
                except ValueError:
                    break
            keys, vals = keys[:keep], vals[:keep]
            if keep:
                eng.set_option(eng.OPT_TREE_FLATTEN, int(rng.integers(0, 3)))
                nt = NativeVerkleTree(kl, ext_width=width)
                cut = int(rng.integers(0, keep + 1))
                nt.insert_many(keys[:cut], vals[:cut])
                if cut:
                    assert (nt.commitment(eng, tree_key) == orc.tree_commit(tree_bases, keys[:cut], vals[:cut], ext_width=width)).all(), ("tree-a", kl, width)
                nt.insert_many(keys[cut:], vals[cut:])
                assert (nt.commitment(eng, tree_key) == orc.tree_commit(tree_bases, keys, vals, ext_width=width)).all(), ("tree-b", kl, width, cut)
                nt.close()
                eng.set_option(eng.OPT_TREE_FLATTEN, 0)
        # IPA commitment proofs (ipa/mod.rs:199-265) and the batched multiproof entry (K proofs, one call == K single calls)
        if it % 3 == 0 and N >= 4:
            key = eng.load_key(bases[:N], q=bases[N], window_bits=8)
            Lc, Rc, tc = eng.ipa_prove_commitment_batch(key, a, C)
            for i in range(Bp):
                eL, eR, etip = orc.ipa_prove_commitment(bases, N, a[i], C[i])
                assert (Lc[i] == eL).all() and (Rc[i] == eR).all() and (tc[i] == etip).all(), ("ipa-commitment", N)
            assert eng.ipa_verify_commitment_batch(key, C, Lc, Rc, tc).all()
            K = int(rng.integers(1, 5))
            m_each = rng.integers(1, 12, K).astype(np.uint64)
            mt = int(m_each.sum())
            fb = orc.rand_fr_buf(rng, mt * N).reshape(mt, N, 32)
            Cb = eng.commit_batch(key, fb)
            zb2 = rng.integers(0, N, mt).astype(np.uint64)
            yb2 = np.stack([fb[i, int(zb2[i])] for i in range(mt)])
            outs = eng.multiproof_prove_batch(key, "ipa", fb, Cb, zb2, yb2, m_each)
            off = 0
            for kk_ in range(K):
                me = int(m_each[kk_])
                exp = orc.multiproof_prove("ipa", bases, N, fb[off:off + me], Cb[off:off + me], zb2[off:off + me], yb2[off:off + me])
                assert all((outs[kk_][k] == exp[k]).all() for k in ("D", "L", "R", "tip", "y")), ("mpbatch", N, K, kk_)
                off += me
            key.free()
        # setup paths: IPA CRS from a random seed, KZG Lagrange SRS from the secret and as the group FFT
        if it % 5 == 0:
            seed = rng.bytes(int(rng.integers(0, 130)))
            num = int(rng.integers(1, 300))
            got_crs, nxt = eng.ipa_crs_generate(seed, num)
            exp_crs, exp_nxt = orc.ipa_crs_gen(seed, num)
            assert (got_crs == exp_crs).all() and nxt == exp_nxt, ("crs", len(seed), num)
            gk = eng.load_key(orc.g1_generator()[None], window_bits=8)
            tau = orc.rand_fr(rng, 1)[0]
            mset = int(rng.integers(1, 70))
            lag = eng.kzg_setup_from_secret(gk, orc.fr_to_buf([tau])[0], mset)
            assert (lag == orc.kzg_setup(mset, tau)).all(), ("kzg-setup-secret", mset)
            assert (eng.kzg_setup(eng.kzg_powers(gk, orc.fr_to_buf([tau])[0], mset)) == lag).all(), ("kzg-setup-fft", mset)
            gk.free()
        # MSM (now and then large enough for the optimistic single-pass scatter and both weighted-sum forms)
        n = int(rng.integers(1, 3000)) if it % 6 else int(rng.integers(4096, 40000))
        cm = int(rng.choice([6, 9, 12, 16])) if it % 6 else int(rng.choice([0, 9, 13, 15, 16, 17]))
        mb = orc.points_walk(k1, k0, n)
        mk = eng.load_key(mb, kind=2, window_bits=cm)
        ms = orc.rand_fr_buf(rng, n)
        ms[rng.random(n) < 0.2] = 0
        assert (eng.msm(mk, ms) == orc.msm(mb, ms, mode="pippenger")).all(), ("msm", n, cm)
        mk.free()
    print(f"fuzz ok: {it} iterations in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
