"""Host-only probe of the verkle tree structure (vkzg_tree_* without a GPU): builds a tree from seeded pairs and prints a JSON
summary (status, pairs inserted, node count, a digest of every lookup, sample paths).  Run by tests/test_abi_and_host.py under
different VKZG_TREE_THREADS to compare the threaded bulk load with the sequential insertion."""
import ctypes, hashlib, json, os, sys
import numpy as np
sys.path.insert(0, os.environ["VKZG_ROOT"])
from verkle_kzg_b200 import _lib
L = _lib.lib()
L.vkzg_tree_nodes.restype = ctypes.c_uint64
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
rng = np.random.default_rng(int(sys.argv[1]))
n, kl, mode = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
keys = rng.integers(0, 256, (n, kl), dtype=np.uint8)
if mode == "narrow":      # few first units, deep shared prefixes
    keys[:, 0] = rng.integers(0, 5, n)
    keys[:, 1] = rng.integers(0, 3, n)
if mode == "dups":        # repeated keys: the later value wins
    keys[n // 2:] = keys[: n - n // 2]
if mode == "panic":       # low-entropy keys: sooner or later Node::insert reaches a case the reference panics on
    keys = rng.integers(0, 3, (n, kl), dtype=np.uint8)
    keys[:, 0] = rng.integers(0, 256, n)
vals = rng.integers(0, 256, (n, 32), dtype=np.uint8)
t = ctypes.c_void_p()
assert L.vkzg_tree_create(ctypes.byref(t), ctypes.c_uint32(kl), ctypes.c_uint32(256)) == 0
done = ctypes.c_uint64(0)
st = L.vkzg_tree_insert(t, p(keys), p(vals), ctypes.c_uint64(n), ctypes.byref(done))
h = hashlib.sha256()
out = np.zeros(32, dtype=np.uint8)
found = 0
probe = np.concatenate([keys, rng.integers(0, 256, (200, kl), dtype=np.uint8)])
for k in probe:
    r = L.vkzg_tree_get(t, p(np.ascontiguousarray(k)), p(out))
    found += r
    h.update(bytes([r]) + (bytes(out) if r else b""))
plen = ctypes.c_uint32(0)
ids = np.zeros(kl, dtype=np.uint32)
units = np.zeros(kl, dtype=np.uint8)
paths = []
for k in keys[:50]:
    r = L.vkzg_tree_path_to_stem(t, p(np.ascontiguousarray(k)), ctypes.byref(plen), p(ids), p(units), None, None)
    paths.append([r, plen.value, units[:plen.value].tolist()])
print(json.dumps({"status": st, "done": done.value, "nodes": L.vkzg_tree_nodes(t), "found": found, "digest": h.hexdigest(), "paths": paths}))
