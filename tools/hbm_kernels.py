"""Achieved HBM bandwidth of the streaming (fold / eval) kernels against the measured copy peak (MEASURED_PEAKS.json:
6550.7 GB/s).  CUDA events on the launching stream, inputs larger than L2.  JSON to stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from verkle_kzg_b200 import Engine  # noqa: E402

PEAK = 6550.7


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    torch.cuda.set_device(0)
    eng = Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    res = {"peak_gbs": PEAK, "kernels": {}}

    def rec(name, ms, bytes_moved, note):
        gbs = bytes_moved / (ms * 1e-3) / 1e9
        res["kernels"][name] = {"ms": round(ms, 4), "algorithmic_gb": round(bytes_moved / 1e9, 3), "achieved_gbs": round(gbs, 1),
                                "frac_of_copy_peak": round(gbs / PEAK, 3), "note": note}

    # L1 / I2: element-wise Fr vector ops on 2^24 elements (512 MiB per operand)
    n = 1 << 24
    a = bench.rand_fr_dev(torch, n, gen)
    b = bench.rand_fr_dev(torch, n, gen)
    out = torch.empty_like(a)
    x = bench.rand_fr_dev(torch, 1, gen).cpu().numpy()[0]
    rec("k_fr_vec add (a + b)", timed(lambda: eng.fr_vector_op_dev("add", a, b, None, n, out)), n * 96, "64 B read + 32 B written per element, no multiplication")
    rec("k_fr_vec axpy (a + x b, the IPA fold)", timed(lambda: eng.fr_vector_op_dev("axpy", a, b, x, n, out)), n * 96, "one Fr multiplication per 96 B")
    rec("k_fr_vec scale (a x)", timed(lambda: eng.fr_vector_op_dev("scale", a, None, x, n, out)), n * 64, "one Fr multiplication per 64 B")
    del a, b, out
    # K1 / E1: quotient and evaluation, width 256, 2^15 openings (256 MiB of rows)
    N, B = 256, 1 << 15
    bases = bench.make_points_dev(torch, eng, N, gen)
    key = eng.load_key_dev(bases, N, window_bits=8)
    f = bench.rand_fr_dev(torch, B * N, gen).reshape(B, N, 32)
    tab = np.stack([np.frombuffer(((v << 256) % bench.R_MOD).to_bytes(32, "little"), dtype=np.uint8) for v in range(N)])
    z_in = torch.from_numpy(tab).cuda()[torch.randint(0, N, (B,), device="cuda", generator=gen)].contiguous()
    z_out = bench.rand_fr_dev(torch, B, gen)
    q = torch.empty((B, N, 32), dtype=torch.uint8, device="cuda")
    y = torch.empty((B, 32), dtype=torch.uint8, device="cuda")
    rec("k_poly quotient, in-domain points (K1)", timed(lambda: eng.quotient_batch_dev(key, f, N, z_in, B, q, y)), B * N * 64,
        "8 KiB read + 8 KiB written per opening, 2 Fr multiplications per element (table of 1/(w^d - 1))")
    rec("k_poly quotient, outside-domain points (K2)", timed(lambda: eng.quotient_batch_dev(key, f, N, z_out, B, q, y)), B * N * 64,
        "+ barycentric evaluation and one shared inversion per opening: ~12 Fr multiplications per element -> integer-bound")
    rec("k_poly evaluate, in-domain points (E1)", timed(lambda: eng.evaluate_batch_dev(key, f, N, z_in, B, y)), B * 64,
        "a lookup: 32 B read + 32 B written per opening (latency-bound, listed for completeness)")
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
