"""Multi-GPU correctness (run under torchrun, one rank per GPU): the point-range sharded MSM and the subtree-sharded tree
give the same canonical results as one GPU doing the whole job.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from verkle_kzg_b200 import Engine, _lib  # noqa: E402
from verkle_kzg_b200.sharding import combine_partial_msm, split_range  # noqa: E402
from verkle_kzg_b200.tree import NativeVerkleTree  # noqa: E402


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Engine(local, stream=torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)                                   # same data on every rank
    n = 1 << 16
    pts = bench.make_points_dev(torch, eng, n, gen)
    s = bench.rand_fr_dev(torch, n, gen)
    # ---- MSM: every rank takes a point range, partial sums are gathered and added
    first, cnt = split_range(n, world, rank)
    key = eng.load_key_dev(pts[first:first + cnt].contiguous(), cnt, kind=_lib.KEY_MSM)
    part = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
    eng.msm_dev(key, s[first:first + cnt].contiguous(), cnt, part)
    total = combine_partial_msm(dist, torch, eng, part)
    full_key = eng.load_key_dev(pts, n, kind=_lib.KEY_MSM)
    full = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
    eng.msm_dev(full_key, s, n, full)
    eng.sync()
    assert torch.equal(total.cpu(), full.cpu()), "sharded MSM differs from the single-GPU MSM"
    # ---- tree: whole subtrees per rank (ranges of the root's child index), partial roots added
    rng = np.random.default_rng(9)
    nk = 1 << 14
    keys = rng.integers(0, 256, (nk, 32), dtype=np.uint8)
    vals = rng.integers(0, 256, (nk, 32), dtype=np.uint8)
    bases = bench.make_points_dev(torch, eng, 256, gen)
    tkey = eng.load_key_dev(bases, 256, window_bits=12)
    lo, c = split_range(256, world, rank)
    sel = (keys[:, 0] >= lo) & (keys[:, 0] < lo + c)
    t = NativeVerkleTree(32, 256)
    t.insert_many(keys[sel], vals[sel])
    mine = torch.from_numpy(t.commitment(eng, tkey)).cuda()
    root = combine_partial_msm(dist, torch, eng, mine)
    tf = NativeVerkleTree(32, 256)
    tf.insert_many(keys, vals)
    full_root = tf.commitment(eng, tkey)
    eng.sync()
    assert (root.cpu().numpy().reshape(64) == full_root).all(), "sharded tree root differs from the single-GPU root"
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check ok on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
