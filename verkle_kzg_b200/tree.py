"""Host side of T1: the verkle tree of verkle-tree/src/{lib,node}.rs as the CALLER of the batched node
commitment — a literal mirror of Node::insert (node.rs:133-197), flattened into the per-level sparse term
lists that vkzg_tree_commit_levels / vkzg_tree_level_dev consume.

Reference behaviour kept as is (the root commitment depends on it):
  * the "stem" of a key is the WHOLE key (lib.rs:62-68), so every distinct key owns one Extension node with a
    single leaf at index key[-1];
  * when a new key meets an Extension with a different stem, ONE Internal node is created whose two children
    are filed under the first differing unit d' (node.rs:166-181), but later inserts index that node by its tree
    depth (node.rs:188-190) — the structure therefore depends on insertion order; it is reproduced, not fixed;
  * extension layout (node.rs:226-240): leaf idx puts its 16-byte halves (low, high) at slots (2 idx) % W and
    (2 idx + 1) % W of C1 if idx < W/2 else C2, with W = key length in the reference (quirk Q6), 256 in the
    Ethereum layout; C = commit([1, from_le_bytes_mod_order(stem), to_data_item(C1), to_data_item(C2)]).
"""
import ctypes

import numpy as np

from . import _lib

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_MONT_R = 1 << 256


def _mont(v):
    return np.frombuffer(((v % R_MOD) * _MONT_R % R_MOD).to_bytes(32, "little"), dtype=np.uint8)


def _fr_mont_from_le_bytes(b):
    return _mont(int.from_bytes(bytes(b), "little"))


class _Ext:
    __slots__ = ("stem", "leaves", "level", "gid")

    def __init__(self, stem):
        self.stem, self.leaves = stem, {}


class _Int:
    __slots__ = ("children", "level", "gid")

    def __init__(self):
        self.children = {}


def _insert(root, key, value):
    """Node::insert (node.rs:133-197)"""
    n = len(key)
    node, depth = root, 0
    while True:
        if depth >= n:
            raise ValueError("internal chain deeper than the key")  # reference: index out of bounds
        k = key[depth]
        child = node.children.get(k)
        if child is None:
            e = _Ext(key)
            e.leaves[key[n - 1]] = value
            node.children[k] = e
            return
        if isinstance(child, _Ext):
            if child.stem == key or depth == n - 2:
                if child.stem != key:
                    raise ValueError("Traversed to extension node with differing stem")  # node.rs:139-141
                child.leaves[key[n - 1]] = value
                return
            d = depth + 1
            while d < n and child.stem[d] == key[d]:  # next_diff_depth, lib.rs:50-59
                d += 1
            if d >= n:  # the stems only differ above this depth: the reference indexes out of bounds (panic)
                raise ValueError("stems differ only above the current depth")
            inner = _Int()
            e = _Ext(key)
            e.leaves[key[n - 1]] = value
            inner.children[key[d]] = e
            inner.children[child.stem[d]] = child
            node.children[k] = inner
            return
        node, depth = child, depth + 1


def _flatten(root, W):
    """post-order walk -> level lists.  Level 0: non-empty C1 / C2 vectors; level 1: extension nodes;
    level >= 2: internal nodes by height; the root alone in the last level."""
    lv = {}  # level -> list of (terms) ; a term is (slot, child_gid or -1, literal or None)
    one = _mont(1)

    def add(level, terms):
        lst = lv.setdefault(level, [])
        lst.append(terms)
        return len(lst) - 1

    # iterative post-order
    handles = {}  # id(node) -> (level, index in level)
    stack = [(root, False)]
    while stack:
        node, done = stack.pop()
        if isinstance(node, _Ext):
            c1, c2 = {}, {}
            for idx, val in node.leaves.items():
                tgt = c1 if idx < W // 2 else c2
                tgt[(2 * idx) % W] = _fr_mont_from_le_bytes(val[:16])
                tgt[(2 * idx + 1) % W] = _fr_mont_from_le_bytes(val[16:])
            terms = [(0, None, one), (1, None, _fr_mont_from_le_bytes(node.stem))]
            for slot, vec in ((2, c1), (3, c2)):
                if vec:
                    h = (0, add(0, [(s, None, v) for s, v in sorted(vec.items())]))
                    terms.append((slot, h, None))
            handles[id(node)] = (1, add(1, terms))
            continue
        if not done:
            stack.append((node, True))
            for ch in node.children.values():
                stack.append((ch, False))
            continue
        terms = []
        level = 2
        for unit, ch in sorted(node.children.items()):
            h = handles[id(ch)]
            level = max(level, h[0] + 1)
            terms.append((unit, h, None))
        handles[id(node)] = (level, add(level, terms))
    # the root must be alone in the last level
    top = max(lv.keys())
    assert len(lv[top]) == 1
    # internal levels: nodes ordered by their number of children (lanes of one warp then walk term lists of similar
    # length); handles of the moved nodes are remapped
    remap = {}
    for L in lv:
        if L >= 2 and len(lv[L]) >= 64:
            order_l = sorted(range(len(lv[L])), key=lambda j: -len(lv[L][j]))
            lv[L] = [lv[L][j] for j in order_l]
            for new, old in enumerate(order_l):
                remap[(L, old)] = (L, new)
    if remap:
        for L in lv:
            lv[L] = [[(slot, remap.get(h, h) if h is not None else None, lit) for slot, h, lit in terms] for terms in lv[L]]
    # global ids: levels in increasing order (empty levels dropped)
    order = sorted(lv.keys())
    base, acc = {}, 0
    for L in order:
        base[L] = acc
        acc += len(lv[L])
    out = []
    zero = np.zeros(32, dtype=np.uint8)
    for L in order:
        nodes = lv[L]
        row_ptr = np.zeros(len(nodes) + 1, dtype=np.uint32)
        slots, childs, lits = [], [], []
        for j, terms in enumerate(nodes):
            for slot, h, lit in terms:
                slots.append(slot)
                if h is None:
                    childs.append(-1)
                    lits.append(lit)
                else:
                    childs.append(base[h[0]] + h[1])
                    lits.append(zero)
            row_ptr[j + 1] = len(slots)
        out.append(dict(row_ptr=row_ptr, slot=np.array(slots, dtype=np.uint16), child=np.array(childs, dtype=np.int32),
                        lit=np.stack(lits) if lits else np.zeros((0, 32), dtype=np.uint8)))
    return out


def build_levels(keys, values, ext_width):
    """insert (key, value) pairs IN ORDER into an empty tree and flatten it.  keys uint8[n, key_len], values uint8[n, 32]."""
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    values = np.ascontiguousarray(values, dtype=np.uint8)
    root = _Int()
    for i in range(len(keys)):
        _insert(root, keys[i].tobytes(), values[i].tobytes())
    return _flatten(root, int(ext_width))


class VerkleTree:
    """Mirror of verkle-tree/src/lib.rs:87-137 for the commitment path: insert_single / get_single / commitment."""

    def __init__(self, key_len, ext_width=None):
        self.key_len = key_len
        self.ext_width = key_len if ext_width is None else ext_width  # reference: const generic N (quirk Q6)
        self.root = _Int()
        self._commit = None

    def insert_single(self, key, value):
        key, value = bytes(key), bytes(value)
        assert len(key) == self.key_len and len(value) == 32
        _insert(self.root, key, value)
        self._commit = None  # node.rs:145,156: cached commitments on the path are cleared

    def get_single(self, key):
        """Node::get_stem + leaf lookup (node.rs:74-93)"""
        key = bytes(key)
        node, depth = self.root, 0
        while isinstance(node, _Int):
            node = node.children.get(key[depth])
            depth += 1
            if node is None:
                return None
        return node.leaves.get(key[-1]) if node.stem == key else None

    def path_to_stem(self, stem):
        """VerkleTree::path_to_stem (lib.rs:131-137, node.rs:101-119): [(prefix bytes, unit, internal node)] for every
        internal node on the way to `stem`; ValueError = VerkleError::InvalidPath"""
        stem = bytes(stem)
        path, node = [], self.root
        while isinstance(node, _Int):
            depth = len(path)
            child = node.children.get(stem[depth])
            if child is None:
                raise ValueError("InvalidPath")
            path.append((stem[:depth + 1], stem[depth], node))
            node = child
        return path

    def levels(self):
        return _flatten(self.root, self.ext_width)

    def commitment(self, engine, key):
        """VerkleTree::commitment (lib.rs:127-129): the root commitment, every node recommitted level by level"""
        if self._commit is None:
            self._commit = engine.tree_commit_levels(key, self.levels())
        return self._commit


class NativeVerkleTree:
    """The same interface over libvkzg's native host tree (vkzg_tree_*): C++ insertion, cached node commitments,
    incremental recommit of the dirty paths.  Use this one for bulk work; `VerkleTree` above is the readable mirror."""

    def __init__(self, key_len, ext_width=None):
        self._L = _lib.lib()
        self._L.vkzg_tree_nodes.restype = ctypes.c_uint64
        self.key_len = key_len
        self.ext_width = key_len if ext_width is None else ext_width
        self._t = ctypes.c_void_p()
        _lib.check(self._L.vkzg_tree_create(ctypes.byref(self._t), ctypes.c_uint32(key_len), ctypes.c_uint32(self.ext_width)),
                   "vkzg_tree_create")
        self.last_committed = 0

    def close(self):
        if self._t:
            self._L.vkzg_tree_destroy(self._t)
            self._t = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def insert_single(self, key, value):
        self.insert_many(np.frombuffer(bytes(key), dtype=np.uint8).reshape(1, -1), np.frombuffer(bytes(value), dtype=np.uint8).reshape(1, 32))

    def insert_many(self, keys, values):
        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        values = np.ascontiguousarray(values, dtype=np.uint8)
        assert keys.ndim == 2 and keys.shape[1] == self.key_len and values.shape == (len(keys), 32)
        done = ctypes.c_uint64(0)
        st = self._L.vkzg_tree_insert(self._t, _lib.hptr(keys), _lib.hptr(values), ctypes.c_uint64(len(keys)), ctypes.byref(done))
        if st == -3:
            raise ValueError(f"Traversed to extension node with differing stem (pair {done.value})")
        _lib.check(st, "vkzg_tree_insert")

    def get_single(self, key):
        out = np.zeros(32, dtype=np.uint8)
        k = np.frombuffer(bytes(key), dtype=np.uint8).copy()
        return bytes(out) if self._L.vkzg_tree_get(self._t, _lib.hptr(k), _lib.hptr(out)) == 1 else None

    @property
    def nodes(self):
        return int(self._L.vkzg_tree_nodes(self._t))

    def path_to_stem(self, stem):
        """vkzg_tree_path_to_stem: [(prefix bytes, unit, node id, cached commitment [64] or None when dirty)];
        ValueError = VerkleError::InvalidPath"""
        stem = bytes(stem)
        assert len(stem) == self.key_len
        s = np.frombuffer(stem, dtype=np.uint8).copy()
        n = ctypes.c_uint32(0)
        ids = np.zeros(self.key_len, dtype=np.uint32)
        units = np.zeros(self.key_len, dtype=np.uint8)
        com = np.zeros((self.key_len, 64), dtype=np.uint8)
        clean = np.zeros(self.key_len, dtype=np.uint8)
        st = self._L.vkzg_tree_path_to_stem(self._t, _lib.hptr(s), ctypes.byref(n), _lib.hptr(ids), _lib.hptr(units), _lib.hptr(com),
                                            _lib.hptr(clean))
        if st == -3:
            raise ValueError("InvalidPath")
        _lib.check(st, "vkzg_tree_path_to_stem")
        return [(stem[:d + 1], int(units[d]), int(ids[d]), com[d].copy() if clean[d] else None) for d in range(n.value)]

    def commitment(self, engine, key):
        out = np.zeros(64, dtype=np.uint8)
        n = ctypes.c_uint64(0)
        _lib.check(self._L.vkzg_tree_commit(engine._ctx, ctypes.c_uint32(key.id), self._t, _lib.hptr(out), ctypes.byref(n)), "vkzg_tree_commit")
        self.last_committed = n.value
        return out
