"""The hot loops' Fq multipliers on the device (csrc/field.cuh): the inline-PTX carry chains of the dedicated square
(fp_sqr_lazy: 36 + 72 instead of 64 + 72 limb products) only exist in device code — tests/test_host_arith.py runs the
host emulation of the same control flow — so they are checked here against Python integers through the C ABI's probe."""
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


def _raw(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).reshape(-1, 32).copy()


def _operands():
    mod = orc.P_MOD
    rng = np.random.default_rng(4242)
    edge = [0, 1, 2, mod - 1, mod, mod + 1, 2 * mod - 1, 2 * mod, (1 << 254) - 1, (1 << 254), mod + (1 << 253), 2 * mod - (1 << 32),
            (1 << 31), (1 << 32) - 1, (1 << 63), ((1 << 224) - 1) << 29]
    pat = [sum(0x80000000 << (32 * i) for i in range(8)), sum(0xFFFFFFFF << (32 * i) for i in range(0, 8, 2)),
           sum(0xFFFFFFFF << (32 * i) for i in range(1, 8, 2)), sum(0x80000001 << (32 * i) for i in range(8)), (1 << 256) - 1]
    edge += [v % (2 * mod + 1) for v in pat] + [min(v, 2 * mod) for v in pat]
    for top in range(0, 0x61):
        v = (top << 248) | ((1 << 248) - 1)
        if v <= 2 * mod:
            edge.append(v)
    return [int.from_bytes(rng.bytes(32), "little") % (2 * mod + 1) for _ in range(20000)] + edge


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("iters", [1, 2, 7])
def test_square_chain_matches_python(eng, mode, iters):
    """x <- x^(2^iters) in Montgomery arithmetic, operands anywhere in [0, 2p] (the lazy range of the accumulation loops)"""
    import torch
    mod = orc.P_MOD
    xs = _operands()
    d = torch.from_numpy(_raw(xs)).cuda()
    eng.probe_fq_sqr_dev(d, len(xs), iters, mode)
    eng.sync()
    got = [int.from_bytes(bytes(r), "little") for r in d.cpu().numpy()]
    rinv = pow(orc.MONT_R, -1, mod)
    want = []
    for x in xs:
        for _ in range(iters):
            x = x * x * rinv % mod
        want.append(x)
    assert got == want


def test_square_equals_product_at_scale(eng):
    """2^20 random operands, 64 dependent squarings: the dedicated square and the general product agree on every element"""
    import torch
    g = torch.Generator().manual_seed(99)
    raw = torch.randint(0, 256, (1 << 20, 32), dtype=torch.uint8, generator=g)
    raw[:, 31] &= 0x3F   # < 2^254 < 2p
    a = raw.cuda()
    b = a.clone()
    eng.probe_fq_sqr_dev(a, a.shape[0], 64, 0)
    eng.probe_fq_sqr_dev(b, b.shape[0], 64, 1)
    eng.sync()
    assert torch.equal(a, b)


@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("iters", [1, 3])
def test_product_chain_matches_python(eng, mode, iters):
    """x <- x b^iters with b = x's halves swapped (both signs of the Karatsuba middle term occur), general (3) / Karatsuba (4)"""
    import torch
    mod = orc.P_MOD
    xs = _operands()
    M128 = (1 << 128) - 1
    # operands whose halves are equal or differ in the lowest bits only, and all-ones halves
    xs += [(h << 128) | h for h in (0, 1, M128 >> 3, 12345)] + [((M128 >> 3) << 128) | M128, (1 << 128) | 2, (2 << 128) | 1]
    d = torch.from_numpy(_raw(xs)).cuda()
    eng.probe_fq_sqr_dev(d, len(xs), iters, mode)
    eng.sync()
    got = [int.from_bytes(bytes(r), "little") for r in d.cpu().numpy()]
    rinv = pow(orc.MONT_R, -1, mod)
    want = []
    for x in xs:
        b = ((x & M128) << 128 | (x >> 128)) & ((1 << 254) - 1)
        for _ in range(iters):
            x = x * b * rinv % mod
        want.append(x)
    assert got == want


def test_karatsuba_equals_general_product_at_scale(eng):
    import torch
    g = torch.Generator().manual_seed(77)
    raw = torch.randint(0, 256, (1 << 20, 32), dtype=torch.uint8, generator=g)
    raw[:, 31] &= 0x3F
    a = raw.cuda()
    b = a.clone()
    eng.probe_fq_sqr_dev(a, a.shape[0], 33, 3)
    eng.probe_fq_sqr_dev(b, b.shape[0], 33, 4)
    eng.sync()
    assert torch.equal(a, b)
