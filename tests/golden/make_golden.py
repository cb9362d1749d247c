"""Generates tests/golden/vectors.npz from the CPU oracle on fixed seeds.

The reference ships no golden vectors and cannot be run here (Rust, no toolchain), so these
fixtures pin the ORACLE's restatement: they make every later change to the oracle or to libvkzg visible as a byte
difference.  Regenerate with:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402


def main():
    rng = np.random.default_rng(0x5EED)
    out = {}
    N = 32
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    out["bases"] = bases
    # commit + IPA opening, in-domain and outside
    a = orc.rand_fr_buf(rng, 2 * N).reshape(2, N, 32)
    out["ipa_a"] = a
    C = orc.commit_batch(bases[:N], a)
    out["ipa_C"] = C
    z = orc.fr_to_buf([5, 3 * N + 1])
    out["ipa_z"] = z
    for i in range(2):
        L, R, tip, y = orc.ipa_prove(bases, N, a[i], C[i], z[i])
        out[f"ipa_L{i}"], out[f"ipa_R{i}"], out[f"ipa_tip{i}"], out[f"ipa_y{i}"] = L, R, tip, y
    # transcript challenges (hash_to_field through the state machine)
    out["h2f_abc_ipa"] = orc.hash_to_fr(b"abc", "ipa")
    out["h2f_long_multiproof"] = orc.hash_to_fr(bytes(range(200)), "multiproof")
    # KZG over the tau = 100 SRS
    srs = orc.kzg_setup(N, 100)
    out["kzg_srs"] = srs
    f = orc.rand_fr_buf(rng, N)
    out["kzg_f"] = f
    for name, pt in (("in", 7), ("out", 2 * N + 3)):
        pf, y, ok = orc.kzg_prove(srs, f, orc.fr_to_buf([pt])[0])
        assert ok
        out[f"kzg_proof_{name}"], out[f"kzg_y_{name}"] = pf, y
    # multiproof (IPA), 12 queries
    m = 12
    fm = orc.rand_fr_buf(rng, m * N).reshape(m, N, 32)
    Cm = orc.commit_batch(bases[:N], fm)
    zm = rng.integers(0, N, m).astype(np.uint64)
    ym = np.stack([fm[i, int(zm[i])] for i in range(m)])
    mp = orc.multiproof_prove("ipa", bases, N, fm, Cm, zm, ym)
    out["mp_f"], out["mp_C"], out["mp_z"], out["mp_y"] = fm, Cm, zm, ym
    for k, v in mp.items():
        out[f"mp_out_{k}"] = v
    # verkle tree root, 50 keys of 32 units over a 256-wide key
    kb0, kb1 = orc.rand_fr(rng, 2)
    tb = orc.points_walk(kb0, kb1, 256)
    keys = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    vals = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    out["tree_bases"], out["tree_keys"], out["tree_vals"] = tb, keys, vals
    out["tree_root_w256"] = orc.tree_commit(tb, keys, vals, ext_width=256)
    out["tree_root_w32"] = orc.tree_commit(tb, keys, vals, ext_width=32)
    out["to_data_item"] = orc.to_data_item(bases[:8])
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
