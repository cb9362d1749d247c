"""Round-2 additions to the committed fixtures (tests/golden/vectors_r2.npz), made from the CPU oracle on fixed seeds like
vectors.npz (same caveat: they pin the oracle's restatement and make drift visible; they do not pin arkworks).
    python tests/golden/make_golden_r2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402


def main():
    rng = np.random.default_rng(0x5EED2)
    out = {}
    # IPA CRS (ipa_point_generator.rs:51-109): default seed and a 70-byte seed (more than one SHA-256 block)
    pts, nxt = orc.ipa_crs_gen(b"eth_verkle_oct_2021", 24)
    out["crs_default"], out["crs_default_next"] = pts, np.array([nxt], dtype=np.uint64)
    seed2 = bytes(range(70))
    pts2, nxt2 = orc.ipa_crs_gen(seed2, 9)
    out["crs_seed70"], out["crs_seed70_next"] = pts2, np.array([nxt2], dtype=np.uint64)
    # IPA commitment proof (ipa/mod.rs:199-265) over the first 17 CRS points (N = 16 + q)
    N = 16
    bases = pts[:N + 1]
    out["cp_bases"] = bases
    a = orc.rand_fr_buf(rng, N)
    C = orc.commit_batch(bases[:N], a[None])[0]
    L, R, tip = orc.ipa_prove_commitment(bases, N, a, C)
    out["cp_a"], out["cp_C"], out["cp_L"], out["cp_R"], out["cp_tip"] = a, C, L, R, tip
    # KZG setup with a random secret and a size that is not a power of two (zero-padded inverse FFT)
    tau = orc.rand_fr(rng, 1)[0]
    out["setup_tau"] = orc.fr_to_buf([tau])[0]
    out["setup_m20"] = orc.kzg_setup(20, tau)
    # one MSM whose window widths exercise the top-window paths (3000 points)
    k0, k1 = orc.rand_fr(rng, 2)
    mb = orc.points_walk(k0, k1, 3000)
    ms = orc.rand_fr_buf(rng, 3000)
    out["msm_k"] = orc.fr_to_buf([k0, k1])
    out["msm_scalars"] = ms
    out["msm_result"] = orc.msm(mb, ms, mode="pippenger")
    np.savez_compressed(os.path.join(HERE, "vectors_r2.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
