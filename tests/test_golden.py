"""Committed fixtures (tests/golden/vectors.npz, made by tests/golden/make_golden.py from the oracle on fixed seeds).
CPU: the oracle still reproduces them.  GPU: libvkzg reproduces them through the C ABI."""
import os

import numpy as np
import pytest

import orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.npz"))
G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors_r2.npz"))  # make_golden_r2.py
N = 32


def test_oracle_reproduces_golden():
    bases = G["bases"]
    assert (orc.commit_batch(bases[:N], G["ipa_a"]) == G["ipa_C"]).all()
    for i in range(2):
        L, R, tip, y = orc.ipa_prove(bases, N, G["ipa_a"][i], G["ipa_C"][i], G["ipa_z"][i])
        assert (L == G[f"ipa_L{i}"]).all() and (R == G[f"ipa_R{i}"]).all() and (tip == G[f"ipa_tip{i}"]).all() and (y == G[f"ipa_y{i}"]).all()
    assert (orc.hash_to_fr(b"abc", "ipa") == G["h2f_abc_ipa"]).all()
    assert (orc.hash_to_fr(bytes(range(200)), "multiproof") == G["h2f_long_multiproof"]).all()
    assert (orc.kzg_setup(N, 100) == G["kzg_srs"]).all()
    mp = orc.multiproof_prove("ipa", bases, N, G["mp_f"], G["mp_C"], G["mp_z"], G["mp_y"])
    for k, v in mp.items():
        assert (v == G[f"mp_out_{k}"]).all(), k
    assert (orc.tree_commit(G["tree_bases"], G["tree_keys"], G["tree_vals"], ext_width=256) == G["tree_root_w256"]).all()
    assert (orc.to_data_item(bases[:8]) == G["to_data_item"]).all()


def test_oracle_reproduces_round2_golden():
    pts, nxt = orc.ipa_crs_gen(b"eth_verkle_oct_2021", 24)
    assert (pts == G2["crs_default"]).all() and nxt == int(G2["crs_default_next"][0])
    pts2, nxt2 = orc.ipa_crs_gen(bytes(range(70)), 9)
    assert (pts2 == G2["crs_seed70"]).all() and nxt2 == int(G2["crs_seed70_next"][0])
    L, R, tip = orc.ipa_prove_commitment(G2["cp_bases"], 16, G2["cp_a"], G2["cp_C"])
    assert (L == G2["cp_L"]).all() and (R == G2["cp_R"]).all() and (tip == G2["cp_tip"]).all()
    assert orc.ipa_verify_commitment(G2["cp_bases"], 16, G2["cp_C"], L, R, tip)
    assert (orc.kzg_setup(20, orc.buf_to_fr(G2["setup_tau"][None])[0]) == G2["setup_m20"]).all()
    k0, k1 = orc.buf_to_fr(G2["msm_k"])
    assert (orc.msm(orc.points_walk(k0, k1, 3000), G2["msm_scalars"], mode="pippenger") == G2["msm_result"]).all()


@pytest.mark.gpu
def test_libvkzg_reproduces_round2_golden():
    from verkle_kzg_b200 import Engine
    eng = Engine(0)
    pts, nxt = eng.ipa_crs_generate(b"eth_verkle_oct_2021", 24)
    assert (pts == G2["crs_default"]).all() and nxt == int(G2["crs_default_next"][0])
    pts2, nxt2 = eng.ipa_crs_generate(bytes(range(70)), 9)
    assert (pts2 == G2["crs_seed70"]).all() and nxt2 == int(G2["crs_seed70_next"][0])
    key = eng.load_key(G2["cp_bases"][:16], q=G2["cp_bases"][16], window_bits=8)
    assert (eng.commit_batch(key, G2["cp_a"][None])[0] == G2["cp_C"]).all()
    L, R, tip = eng.ipa_prove_commitment_batch(key, G2["cp_a"][None], G2["cp_C"][None])
    assert (L[0] == G2["cp_L"]).all() and (R[0] == G2["cp_R"]).all() and (tip[0] == G2["cp_tip"]).all()
    assert eng.ipa_verify_commitment_batch(key, G2["cp_C"][None], L, R, tip).all()
    gk = eng.load_key(orc.g1_generator()[None], window_bits=8)
    assert (eng.kzg_setup_from_secret(gk, G2["setup_tau"], 20) == G2["setup_m20"]).all()
    assert (eng.kzg_setup(eng.kzg_powers(gk, G2["setup_tau"], 20)) == G2["setup_m20"]).all()
    k0, k1 = orc.buf_to_fr(G2["msm_k"])
    mb = orc.points_walk(k0, k1, 3000)
    for c in (0, 9, 15, 17):
        mk = eng.load_key(mb, kind=2, window_bits=c)
        assert (eng.msm(mk, G2["msm_scalars"]) == G2["msm_result"]).all(), c
        mk.free()
    eng.close()


@pytest.mark.gpu
def test_libvkzg_reproduces_golden():
    from verkle_kzg_b200 import Engine
    from verkle_kzg_b200.tree import VerkleTree
    eng = Engine(0)
    bases = G["bases"]
    key = eng.load_key(bases[:N], q=bases[N], window_bits=8)
    assert (eng.commit_batch(key, G["ipa_a"]) == G["ipa_C"]).all()
    L, R, tip, y = eng.ipa_prove_batch(key, G["ipa_a"], G["ipa_z"], G["ipa_C"])
    for i in range(2):
        assert (L[i] == G[f"ipa_L{i}"]).all() and (R[i] == G[f"ipa_R{i}"]).all() and (tip[i] == G[f"ipa_tip{i}"]).all() and (y[i] == G[f"ipa_y{i}"]).all()
    mp = eng.multiproof_prove(key, "ipa", G["mp_f"], G["mp_C"], G["mp_z"], G["mp_y"])
    for k in ("D", "L", "R", "tip", "y"):
        assert (mp[k] == G[f"mp_out_{k}"]).all(), k
    assert (eng.to_data_item(bases[:8]) == G["to_data_item"]).all()
    kk = eng.load_key(G["kzg_srs"], window_bits=8)
    for name, pt in (("in", 7), ("out", 2 * N + 3)):
        pf, yy = eng.kzg_open_batch(kk, G["kzg_f"].reshape(1, N, 32), orc.fr_to_buf([pt]))
        assert (pf[0] == G[f"kzg_proof_{name}"]).all() and (yy[0] == G[f"kzg_y_{name}"]).all()
    tk = eng.load_key(G["tree_bases"], window_bits=8)
    for w in (256, 32):
        t = VerkleTree(32, ext_width=w)
        for k, v in zip(G["tree_keys"], G["tree_vals"]):
            t.insert_single(k, v)
        assert (t.commitment(eng, tk) == G[f"tree_root_w{w}"]).all()
    eng.close()
