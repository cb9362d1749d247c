"""T1 parity on the GPU: root commitments equal the oracle's literal Node::insert + gen_commitment."""
import numpy as np
import pytest

import orc


def _reference_safe(keys, vals, key_len):
    """drop the (key, value) pairs on which the reference's Node::insert panics (index out of bounds / differing stem)"""
    from verkle_kzg_b200.tree import VerkleTree
    t = VerkleTree(key_len, 256)
    keep = []
    for i, (k, v) in enumerate(zip(keys, vals)):
        try:
            t.insert_single(k, v)
            keep.append(i)
        except ValueError:
            break  # a failed insert may leave cleared caches behind; stop at the first panic like the reference would
    return keys[keep], vals[keep]

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
    keys, vals = _reference_safe(keys, vals, key_len)
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


@pytest.mark.parametrize("n,key_len,width,low_entropy", [(1, 32, 256, False), (300, 32, 256, True), (500, 4, 256, True), (200, 32, 32, False),
                                                          (150, 40, 256, True),    # stems longer than 32 bytes: host-reduced literals
                                                          (120, 32, 1, False), (120, 32, 2, False), (120, 32, 5, False)])  # narrow extension layouts
@pytest.mark.parametrize("flatten", [0, 1, 2])   # automatic / bulk pass / depth-first walk of the dirty paths
def test_native_tree_incremental_commit(eng, key256, n, key_len, width, low_entropy, flatten):
    """libvkzg's native host tree: full commit, then incremental recommits of only the dirty paths, all equal to the oracle"""
    from verkle_kzg_b200.tree import NativeVerkleTree
    eng.set_option(eng.OPT_TREE_FLATTEN, flatten)
    bases, key = key256
    rng = np.random.default_rng(n * 3 + key_len)
    hi = 4 if low_entropy else 256
    keys = rng.integers(0, hi, (n + 40, key_len), dtype=np.uint8)
    _, first = np.unique(keys[:, : key_len - 1], axis=0, return_index=True)
    keys = keys[np.sort(first)]
    vals = rng.integers(0, 256, (len(keys), 32), dtype=np.uint8)
    keys, vals = _reference_safe(keys, vals, key_len)
    m = max(1, len(keys) - 20)
    t = NativeVerkleTree(key_len, ext_width=width)
    t.insert_many(keys[:m], vals[:m])
    root = t.commitment(eng, key)
    assert (root == orc.tree_commit(bases, keys[:m], vals[:m], ext_width=width)).all()
    full = t.last_committed
    assert (t.commitment(eng, key) == root).all() and t.last_committed == 0     # everything cached
    # insert the remaining keys and overwrite one value: only the touched paths are recommitted
    t.insert_many(keys[m:], vals[m:])
    vals2 = vals.copy()
    vals2[0] = rng.integers(0, 256, 32, dtype=np.uint8)
    t.insert_single(keys[0], vals2[0])
    root2 = t.commitment(eng, key)
    # the oracle replays the SAME insertion sequence (in the reference's order-dependent structure a re-insert of a key
    # that became unreachable creates a second extension instead of overwriting)
    seq_k = np.concatenate([keys, keys[:1]])
    seq_v = np.concatenate([vals, vals2[:1]])
    assert (root2 == orc.tree_commit(bases, seq_k, seq_v, ext_width=width)).all()
    if m > 100:
        assert t.last_committed < full
    t.close()
    eng.set_option(eng.OPT_TREE_FLATTEN, 0)
