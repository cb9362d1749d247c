#!/usr/bin/env python
"""One command to pin verkle_kzg_b200's parity to the REAL reference (arkworks 0.4):

    python tests/golden/arkworks/make_vectors.py /path/to/verkle-kzg      # needs cargo + the crates (network or vendored)

It never touches the reference checkout: both crates are copied to a scratch directory, where
  * `arkworks_vectors.rs` becomes `vector-commit/src/arkworks_vectors.rs`, declared `#[cfg(test)] mod arkworks_vectors;`;
  * the private fields / constructors the vectors need are widened to `pub(crate)` (textual patches listed in PATCHES —
    visibility only, no behaviour);
  * `tree_vector.rs.inc` is inserted into the `mod tests` of `verkle-tree/src/lib.rs` (it uses that module's test types);
then `cargo test ... -- --nocapture` runs the two generator tests and the JSON they print between the ARKVEC markers is
merged into tests/golden/arkworks/vectors.json.  tests/test_arkworks_vectors.py picks that file up: with it, the oracle
(CPU) and libvkzg (GPU) are both checked against arkworks byte for byte; without it those tests SKIP with
"parity unpinned".  This image has no Rust toolchain, so the file is not committed yet.
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))

# (file, pattern, replacement): visibility only
PATCHES = [
    ("vector-commit/src/ipa/mod.rs", r"pub struct IPAProof<G: Group> \{\n    l:", "pub struct IPAProof<G: Group> {\n    pub(crate) l:"),
    ("vector-commit/src/ipa/mod.rs", r"(pub struct IPAProof<G: Group> \{\n    pub\(crate\) l: Vec<G>,\n)    r: Vec<G>,\n    tip: G::ScalarField,\n    y: G::ScalarField,",
     r"\1    pub(crate) r: Vec<G>,\n    pub(crate) tip: G::ScalarField,\n    pub(crate) y: G::ScalarField,"),
    ("vector-commit/src/ipa/mod.rs", r"pub struct IPACommitProof<G: Group> \{\n    l: Vec<G>,\n    r: Vec<G>,\n    tip: G::ScalarField,",
     "pub struct IPACommitProof<G: Group> {\n    pub(crate) l: Vec<G>,\n    pub(crate) r: Vec<G>,\n    pub(crate) tip: G::ScalarField,"),
    ("vector-commit/src/ipa/mod.rs", r"    fn new_from_vec\(all: Vec<G>\)", "    pub(crate) fn new_from_vec(all: Vec<G>)"),
    ("vector-commit/src/ipa/mod.rs", r"\nmod ipa_point_generator;", "\npub(crate) mod ipa_point_generator;"),
    ("vector-commit/src/kzg/mod.rs", r"pub struct KZGProof<F: Field, G: Group> \{\n    proof: KZGCommitment<G>,\n    y: F,",
     "pub struct KZGProof<F: Field, G: Group> {\n    pub(crate) proof: KZGCommitment<G>,\n    pub(crate) y: F,"),
    ("vector-commit/src/kzg/mod.rs", r"\n    lagrange_commitments: Vec<G1>,", "\n    pub(crate) lagrange_commitments: Vec<G1>,"),
    ("vector-commit/src/multiproof.rs", r"pub struct Multiproof<P, D> \{\n    proof: P,\n    d: D,",
     "pub struct Multiproof<P, D> {\n    pub(crate) proof: P,\n    pub(crate) d: D,"),
]


def patch(root):
    for rel, pat, rep in PATCHES:
        p = os.path.join(root, rel)
        s = open(p).read()
        s2, n = re.subn(pat, rep, s, count=1)
        if n != 1:
            raise SystemExit(f"patch did not apply to {rel}: {pat!r} (has the reference changed?)")
        open(p, "w").write(s2)
    shutil.copy(os.path.join(HERE, "arkworks_vectors.rs"), os.path.join(root, "vector-commit/src/arkworks_vectors.rs"))
    lib = os.path.join(root, "vector-commit/src/lib.rs")
    with open(lib, "a") as f:
        f.write("\n#[cfg(test)]\nmod arkworks_vectors;\n")
    tl = os.path.join(root, "verkle-tree/src/lib.rs")
    s = open(tl).read()
    marker = "    #[test]\n    fn test_commitment() {"
    if marker not in s:
        raise SystemExit("verkle-tree/src/lib.rs: test_commitment not found (has the reference changed?)")
    s = s.replace(marker, open(os.path.join(HERE, "tree_vector.rs.inc")).read() + "\n" + marker, 1)
    open(tl, "w").write(s)


def extract(text, begin, end):
    m = re.search(begin + r"\s*\n(.*?)\n\s*" + end, text, re.S)
    if not m:
        raise SystemExit(f"no {begin} block in the cargo output")
    return json.loads(m.group(1))


def main():
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    src = os.path.abspath(sys.argv[1])
    tmp = tempfile.mkdtemp(prefix="arkvec_")
    for crate in ("vector-commit", "verkle-tree"):
        shutil.copytree(os.path.join(src, crate), os.path.join(tmp, crate), ignore=shutil.ignore_patterns("target"))
    for f in ("Cargo.toml", "Cargo.lock"):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), tmp)
    patch(tmp)
    out = {}
    r1 = subprocess.run(["cargo", "test", "--release", "arkworks_vectors", "--", "--nocapture", "--test-threads", "1"],
                        cwd=os.path.join(tmp, "vector-commit"), capture_output=True, text=True)
    if r1.returncode != 0:
        raise SystemExit(r1.stdout + r1.stderr)
    out.update(extract(r1.stdout, "ARKVEC_BEGIN", "ARKVEC_END"))
    r2 = subprocess.run(["cargo", "test", "--release", "arkworks_tree_vector", "--", "--nocapture", "--test-threads", "1"],
                        cwd=os.path.join(tmp, "verkle-tree"), capture_output=True, text=True)
    if r2.returncode != 0:
        raise SystemExit(r2.stdout + r2.stderr)
    out.update(extract(r2.stdout, "ARKVEC_TREE_BEGIN", "ARKVEC_TREE_END"))
    dst = os.path.join(HERE, "vectors.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {dst} ({len(out)} sections); now run: python -m pytest tests/test_arkworks_vectors.py -q")
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
