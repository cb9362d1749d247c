"""Multi-GPU paths, both forms:

 * ONE process, several devices, behind the C ABI (vkzg_mgpu_*: point-range sharded MSM combined over peer copies, batches
   of commits / IPA proofs cut into ranges) — byte-identical to the single-device calls.  Runs on a single GPU too: a group
   may hold several contexts on the same device, which exercises the same sharding and combine code;
 * one process per GPU under torch.distributed.run over NCCL (tools/check_multi_gpu.py: sharded MSM + all_gather +
   vkzg_g1_sum_dev, subtree-sharded tree) — needs >= 2 visible GPUs, skipped otherwise; its output is kept in gpurun_out/."""
import os
import subprocess
import sys

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _groups():
    n = _ngpu()
    groups = [[0], [0, 0, 0]]
    if n >= 2:
        groups.append(list(range(min(n, 8))))
    return groups


@pytest.mark.parametrize("gi", [0, 1, 2])
def test_group_matches_single_device(gi):
    from verkle_kzg_b200 import Engine, MultiEngine
    groups = _groups()
    if gi >= len(groups):
        pytest.skip("needs >= 2 visible GPUs")
    devs = groups[gi]
    rng = np.random.default_rng(100 + gi)
    mg = MultiEngine(devs)
    eng = Engine(0)
    assert mg.size == len(devs)
    # ---- point-range sharded MSM == one-device MSM == oracle (a size that does not divide evenly)
    n = 3001
    k0, k1 = orc.rand_fr(rng, 2)
    pts = orc.points_walk(k0, k1, n)
    s = orc.rand_fr_buf(rng, n)
    mkey = mg.load_key(pts, kind=2, window_bits=9)
    skey = eng.load_key(pts, kind=2, window_bits=9)
    exp = orc.msm(pts, s, mode="pippenger")
    assert (mg.msm(mkey, s) == exp).all() and (eng.msm(skey, s) == exp).all()
    assert (mg.msm(mkey, s[:1000]) == orc.msm(pts[:1000], s[:1000], mode="pippenger")).all()   # zip truncation across slices
    assert (mg.msm(mkey, s[:1]) == orc.msm(pts[:1], s[:1])).all()
    assert not mg.msm(mkey, np.zeros((n, 32), dtype=np.uint8)).any()
    mg.free_key(mkey)
    skey.free()
    # ---- batches: commits and IPA commit + prove, ranges per device
    N = 32
    bases = orc.points_walk(k1, k0, N + 1)
    wkey = mg.load_key(bases[:N], q=bases[N], window_bits=10)
    for B in (1, 2, 7, 64):
        a = orc.rand_fr_buf(rng, B * N).reshape(B, N, 32)
        assert (mg.commit_batch(wkey, a) == orc.commit_batch(bases[:N], a)).all()
        zb = orc.fr_to_buf([int(v) for v in rng.integers(0, 2 * N, B)])
        C, L, R, tip, y = mg.ipa_commit_prove_batch(wkey, a, zb)
        for i in range(0, B, 5):
            eL, eR, etip, ey = orc.ipa_prove(bases, N, a[i], C[i], zb[i])
            assert (L[i] == eL).all() and (R[i] == eR).all() and (tip[i] == etip).all() and (y[i] == ey).all(), (B, i)
    mg.free_key(wkey)
    mg.close()
    eng.close()


def test_one_process_per_gpu_over_nccl():
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 visible GPUs")
    world = 2 if n < 4 else 4
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    with open(os.path.join(out_dir, "check_multi_gpu.log"), "w") as f:
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"multi-GPU check ok on {world} GPUs" in r.stdout
