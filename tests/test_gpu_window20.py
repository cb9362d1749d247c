"""The configuration bench.py measures: the 257-base width-256 key (256 bases + Q) at window_bits = 20 — a 112 GB table
(257 rows-of-bases x 13 windows x 2^19 entries) whose packed (entry index | sign << 31) words reach 1.75e9 of the 2^31
available (commit.cu: emit_entries).  Everything here is byte-for-byte against the oracle:
   * table spot checks: one-hot scalars that select the FIRST and LAST entry of rows of the first / last base, of the
     last window that can borrow (w = 11) and of the top window (w = 12), with either sign;
   * commits at widths 1 / 129 / 256;
   * IPA proofs at in-domain, boundary (N - 1, N, N + 1) and random points, both Q-row paths (prove + device verify);
   * KZG openings (quotient + commit) through the same table.
Skipped with a reason when the device has less than 125 GB free (the table does not fit)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu

C_BITS = 20
W = 13
N = 256


@pytest.fixture(scope="module")
def setup20():
    import torch
    from verkle_kzg_b200 import Engine
    free, _total = torch.cuda.mem_get_info()
    need = 257 * W * (1 << (C_BITS - 1)) * 64 + (6 << 30)
    if free < need:
        pytest.skip(f"window_bits=20 needs {need / 1e9:.0f} GB of free HBM, device has {free / 1e9:.0f} GB free")
    eng = Engine(0)
    rng = np.random.default_rng(0x20C0FFEE)
    k0, k1 = orc.rand_fr(rng, 2)
    bases = orc.points_walk(k0, k1, N + 1)
    key = eng.load_key(bases[:N], q=bases[N], window_bits=C_BITS)
    assert key.table_bytes >= 257 * W * (1 << (C_BITS - 1)) * 64
    yield eng, rng, bases, key
    key.free()
    eng.close()


def _one_hot(pos, value):
    v = [0] * N
    v[pos] = value % orc.R_MOD
    return orc.fr_to_buf(v)


def test_table_spot_checks(setup20):
    """one-hot scalars whose signed digits select a single known table entry (or two adjacent ones)"""
    eng, rng, bases, key = setup20
    half = 1 << (C_BITS - 1)
    rows = []
    for base in (0, 1, 128, N - 1):
        for w in (0, 1, 11):
            rows.append(_one_hot(base, 1 << (C_BITS * w)))                 # first entry (m = 1) of row (base, w)
            rows.append(_one_hot(base, (half - 1) << (C_BITS * w)))        # m = 2^19 - 1, positive
            rows.append(_one_hot(base, half << (C_BITS * w)))              # digit 2^19 -> LAST entry, negative, carry into w + 1
            rows.append(_one_hot(base, (half + 1) << (C_BITS * w)))        # m = 2^19 - 1, negative, carry
        rows.append(_one_hot(base, 1 << (C_BITS * 12)))                    # top window, first entry
        rows.append(_one_hot(base, orc.R_MOD - 1))                         # top digit of r - 1 (largest the top window sees)
        rows.append(_one_hot(base, (orc.R_MOD >> 240) << 240))             # top window alone at its maximum
        rows.append(_one_hot(base, (1 << 240) - 1))                        # all lower windows at 2^20 - 1: a borrow chain into the top
    s = np.stack(rows)
    got = eng.commit_batch(key, s)
    exp = orc.commit_batch(bases[:N], s)
    bad = [i for i in range(len(rows)) if not (got[i] == exp[i]).all()]
    assert not bad, f"table entries disagree with the oracle for one-hot rows {bad}"


@pytest.mark.parametrize("width", [1, 129, 256])
def test_commit_widths(setup20, width):
    eng, rng, bases, key = setup20
    B = 6
    s = orc.rand_fr_buf(rng, B * width).reshape(B, width, 32)
    assert (eng.commit_batch(key, s) == orc.commit_batch(bases[:width], s)).all()


def test_ipa_proofs_match_oracle_and_verify(setup20):
    eng, rng, bases, key = setup20
    zs = [0, 17, N - 1, N, N + 1, 2 * N, orc.R_MOD - 2] + orc.rand_fr(rng, 3)
    B = len(zs)
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    zb = orc.fr_to_buf(zs)
    C, L, R, tip, y = eng.ipa_commit_prove_batch(key, a, zb)
    assert (C == orc.commit_batch(bases[:N], a)).all()
    for i in range(B):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all(), f"proof {i} (z = {zs[i]})"
        assert (tip[i] == etip).all() and (y[i] == ey).all(), f"proof {i} (z = {zs[i]})"
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
    tip2 = tip.copy()
    tip2[4] = tip[5]
    ok = eng.ipa_verify_batch(key, zb, C, L, R, tip2, y)
    assert not ok[4] and ok.sum() == B - 1


def test_ipa_batch_both_half_batches(setup20):
    """a batch large enough for the two-stream half-batch split (>= 8192): all verify, samples of both halves equal the oracle"""
    eng, rng, bases, key = setup20
    B = 8192
    a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    zb = orc.fr_to_buf([int(v) for v in rng.integers(0, N, B)])
    C, L, R, tip, y = eng.ipa_commit_prove_batch(key, a, zb)
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
    for i in (0, 4095, 4096, B - 1):
        eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
        assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all(), i


def test_kzg_open(setup20):
    eng, rng, bases, key = setup20
    B = 5
    f = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
    zb = orc.fr_to_buf([0, 3, N - 1, N + 7] + orc.rand_fr(rng, 1))
    proof, y = eng.kzg_open_batch(key, f, zb, domain_n=N)
    for i in range(B):
        epf, ey, ok = orc.kzg_prove(bases[:N], f[i], zb[i])   # the oracle's prove_point over the key's domain
        assert ok and (proof[i] == epf).all() and (y[i] == ey).all(), i
