"""One-off setup paths timed on the GPU (the reference benches KZG setup at 2048 ... 16384 points,
benches/kzg.rs:45-59):  KZG::setup = group inverse FFT of the powers of tau (vkzg_kzg_setup), the powers themselves
(vkzg_kzg_powers) and IPAPointGenerator::gen (vkzg_ipa_crs_generate).  Each result is checked: the Lagrange SRS against
the closed form L_j(tau) * G computed with scalars (one fixed-base batch), the CRS against the oracle on a prefix.
    python tools/setup_bench.py [--json gpurun_out/setup_bench.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orc  # noqa: E402  (checker only)
from verkle_kzg_b200 import Engine  # noqa: E402

R = orc.R_MOD


def lagrange_scalars(n, tau):
    """L_j(tau) over the radix-2 domain of size n (closed form: (tau^n - 1) w^j / (n (tau - w^j)))"""
    w = pow(5, (R - 1) // n, R)
    num = (pow(tau, n, R) - 1) * pow(n, -1, R) % R
    out, wj = [], 1
    for _ in range(n):
        out.append(num * wj % R * pow((tau - wj) % R, -1, R) % R)
        wj = wj * w % R
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--sizes", default="2048,4096,8192,16384")
    args = ap.parse_args()
    eng = Engine(0)
    g = orc.g1_generator()
    gen_key = eng.load_key(g[None], window_bits=16)
    tau = 0x1234567890ABCDEF1234567890ABCDEF % R
    res = {"kzg_setup": [], "ipa_crs": []}
    for n in [int(x) for x in args.sizes.split(",")]:
        taub = orc.fr_to_buf([tau])[0]
        eng.kzg_powers(gen_key, taub, n)
        t0 = time.perf_counter()
        powers = eng.kzg_powers(gen_key, taub, n)
        t1 = time.perf_counter()
        eng.kzg_setup(powers)
        ts = []
        for _ in range(3):
            t2 = time.perf_counter()
            lag = eng.kzg_setup(powers)
            ts.append(time.perf_counter() - t2)
        want = eng.commit_batch(gen_key, orc.fr_to_buf(lagrange_scalars(n, tau)).reshape(n, 1, 32))
        ok = bool((lag == want).all())
        eng.kzg_setup_from_secret(gen_key, taub, n)
        tc = []
        for _ in range(3):
            t3 = time.perf_counter()
            lag2 = eng.kzg_setup_from_secret(gen_key, taub, n)
            tc.append(time.perf_counter() - t3)
        ok = ok and bool((lag2 == want).all())
        res["kzg_setup"].append({"n": n, "powers_ms": (t1 - t0) * 1e3, "setup_ms": min(ts) * 1e3, "points_per_s": n / min(ts),
                                 "from_secret_ms": min(tc) * 1e3, "ok": ok})
        print(res["kzg_setup"][-1], flush=True)
        assert ok
    for num in (256, 4096, 65536, 1 << 20):
        eng.ipa_crs_generate(b"eth_verkle_oct_2021", num)
        t0 = time.perf_counter()
        pts, nxt = eng.ipa_crs_generate(b"eth_verkle_oct_2021", num)
        dt = time.perf_counter() - t0
        k = min(num, 256)
        t1 = time.perf_counter()
        want, _ = orc.ipa_crs_gen(b"eth_verkle_oct_2021", k)
        cpu = time.perf_counter() - t1
        ok = bool((pts[:k] == want).all())
        res["ipa_crs"].append({"num": num, "candidates": nxt, "ms": dt * 1e3, "points_per_s": num / dt, "cpu_points_per_s_1thread": k / cpu, "ok": ok})
        print(res["ipa_crs"][-1], flush=True)
        assert ok
    if args.json:
        json.dump(res, open(args.json, "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    main()
