"""Single-call latencies (B = 1) and verifier throughput through the host-pointer C ABI, width 256.  JSON to stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic CRS / vectors made on the device by the product itself)
from verkle_kzg_b200 import Engine  # noqa: E402

R_MOD = bench.R_MOD


def fr(vals):
    """F::from(v) in the ABI layout (Montgomery)"""
    return np.stack([np.frombuffer(((int(v) << 256) % R_MOD).to_bytes(32, "little"), dtype=np.uint8) for v in vals])


def timeit(fn, reps=20):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    eng = Engine(0)
    rng = np.random.default_rng(1)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    N = 256
    bases = bench.make_points_dev(torch, eng, N + 1, gen).cpu().numpy()
    key = eng.load_key(bases[:N], q=bases[N], window_bits=16)
    a1 = bench.rand_fr_dev(torch, N, gen).cpu().numpy().reshape(1, N, 32)
    z_in, z_out = fr([7]), bench.rand_fr_dev(torch, 1, gen).cpu().numpy()
    C1 = eng.commit_batch(key, a1)
    res = {"unit": "ms per call, B = 1, width 256, c = 16, host pointers"}
    res["commit"] = timeit(lambda: eng.commit_batch(key, a1))
    res["ipa_prove_in_domain"] = timeit(lambda: eng.ipa_prove_batch(key, a1, z_in, C1))
    res["ipa_prove_out_domain"] = timeit(lambda: eng.ipa_prove_batch(key, a1, z_out, C1))
    L, R, tip, y = eng.ipa_prove_batch(key, a1, z_in, C1)
    res["ipa_verify"] = timeit(lambda: eng.ipa_verify_batch(key, z_in, C1, L, R, tip, y))
    res["kzg_open_in_domain"] = timeit(lambda: eng.kzg_open_batch(key, a1, z_in))
    res["kzg_open_out_domain"] = timeit(lambda: eng.kzg_open_batch(key, a1, z_out))
    # batched verifier throughput
    B = 4096
    a = bench.rand_fr_dev(torch, B * N, gen).cpu().numpy().reshape(B, N, 32)
    zb = fr(rng.integers(0, N, B))
    C = eng.commit_batch(key, a)
    L, R, tip, y = eng.ipa_prove_batch(key, a, zb, C)
    ms = timeit(lambda: eng.ipa_verify_batch(key, zb, C, L, R, tip, y), reps=5)
    assert eng.ipa_verify_batch(key, zb, C, L, R, tip, y).all()
    res["ipa_verify_batch_4096_ms"] = ms
    res["ipa_verifies_per_s"] = B / (ms * 1e-3)
    # multiproof prove + verify, 4096 queries
    zq = rng.integers(0, N, B).astype(np.uint64)
    yq = np.stack([a[i, int(zq[i])] for i in range(B)])
    mp = eng.multiproof_prove(key, "ipa", a, C, zq, yq)
    res["multiproof_prove_4096_ms"] = timeit(lambda: eng.multiproof_prove(key, "ipa", a, C, zq, yq), reps=5)
    res["multiproof_verify_4096_ms"] = timeit(lambda: eng.multiproof_verify_ipa(key, C, zq, yq, mp), reps=5)
    assert eng.multiproof_verify_ipa(key, C, zq, yq, mp)
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}, indent=1))


if __name__ == "__main__":
    main()
