"""Committed fixtures (tests/golden/vectors.npz, made by tests/golden/make_golden.py from the oracle on fixed seeds).
CPU: the oracle still reproduces them.  GPU: libvkzg reproduces them through the C ABI."""
import os

import numpy as np
import pytest

import orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.npz"))
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
