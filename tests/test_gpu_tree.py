"""T1 parity on the GPU: root commitments equal the oracle's literal Node::insert + gen_commitment."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from verkle_kzg_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def key256(eng):
    rng = np.random.default_rng(600)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, 256)
    return bases, eng.load_key(bases)


@pytest.mark.parametrize("n,key_len,width,low_entropy", [(1, 32, 256, False), (40, 32, 256, False), (300, 32, 256, True),
                                                          (60, 3, 3, False), (500, 4, 256, True), (200, 32, 32, False)])
def test_tree_root_matches_oracle(eng, key256, n, key_len, width, low_entropy):
    from verkle_kzg_b200.tree import VerkleTree
    bases, key = key256
    rng = np.random.default_rng(n + key_len)
    hi = 4 if low_entropy else 256  # few distinct units -> deep shared prefixes, path-compressed internals
    keys = rng.integers(0, hi, (n, key_len), dtype=np.uint8)
    # the reference panics when two keys differ only in their last unit: make the first key_len-1 units distinct
    _, first = np.unique(keys[:, : key_len - 1], axis=0, return_index=True)
    keys = keys[np.sort(first)]
    vals = rng.integers(0, 256, (len(keys), 32), dtype=np.uint8)
    t = VerkleTree(key_len, ext_width=width)
    for k, v in zip(keys, vals):
        t.insert_single(k, v)
    root = t.commitment(eng, key)
    assert (root == orc.tree_commit(bases, keys, vals, ext_width=width)).all()
    assert t.get_single(keys[0]) == bytes(vals[0])
    # overwrite a value (verkle-tree/src/lib.rs:305-317) and recommit
    vals2 = vals.copy()
    vals2[0] = rng.integers(0, 256, 32, dtype=np.uint8)
    t.insert_single(keys[0], vals2[0])
    root2 = t.commitment(eng, key)
    assert (root2 == orc.tree_commit(bases, keys, vals2, ext_width=width)).all()
    assert (root2 != root).any()


def test_empty_tree(eng, key256):
    from verkle_kzg_b200.tree import VerkleTree
    _, key = key256
    assert not VerkleTree(32, 256).commitment(eng, key).any()
