"""Two multiproofs over 2^12 openings (width 256) through the host-pointer C ABI — the subject of an ncu launch list that
shows how the 2.9 ms of configs[2] split between the streaming part and the single inner IPA opening."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from verkle_kzg_b200 import Engine  # noqa: E402


def main():
    eng = Engine(0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    rng = np.random.default_rng(1)
    N, m = 256, 4096
    bases = bench.make_points_dev(torch, eng, N + 1, gen).cpu().numpy()
    key = eng.load_key(bases[:N], q=bases[N], window_bits=int(os.environ.get("VKZG_TRACE_WINDOW_BITS", "16")))
    f = bench.rand_fr_dev(torch, m * N, gen).cpu().numpy().reshape(m, N, 32)
    C = eng.commit_batch(key, f)
    z = rng.integers(0, N, m).astype(np.uint64)
    y = np.stack([f[i, int(z[i])] for i in range(m)])
    for _ in range(2):
        eng.multiproof_prove(key, "ipa", f, C, z, y)
    print("launches", eng.launches)


if __name__ == "__main__":
    main()
