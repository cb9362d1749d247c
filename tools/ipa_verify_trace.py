"""Three single (B = 1) width-256 IPA verifications through the host-pointer C ABI — subject of an ncu launch list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from verkle_kzg_b200 import Engine  # noqa: E402


def main():
    eng = Engine(0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    N = 256
    bases = bench.make_points_dev(torch, eng, N + 1, gen).cpu().numpy()
    key = eng.load_key(bases[:N], q=bases[N], window_bits=int(os.environ.get("VKZG_TRACE_WINDOW_BITS", "16")))
    a1 = bench.rand_fr_dev(torch, N, gen).cpu().numpy().reshape(1, N, 32)
    z = bench.rand_fr_dev(torch, 1, gen).cpu().numpy()
    C1 = eng.commit_batch(key, a1)
    L, R, tip, y = eng.ipa_prove_batch(key, a1, z, C1)
    for _ in range(3):
        assert eng.ipa_verify_batch(key, z, C1, L, R, tip, y).all()
    print("launches", eng.launches)


if __name__ == "__main__":
    main()
