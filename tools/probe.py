"""Integer-pipe probes on the B200: dependent-free multiply-accumulate chains (the roofline denominator
for the MSM / commit / IPA kernels — MEASURED_PEAKS.json has no integer figure) and the Fq multiplier's
throughput.  Writes one JSON document to stdout.

    python tools/probe.py > gpurun_out/probe.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from verkle_kzg_b200 import Engine  # noqa: E402


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best


def main():
    torch.cuda.set_device(0)
    eng = Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    res = {"gpu": torch.cuda.get_device_name(0), "sms": sm, "imad": {}, "fq_mul": {}}
    names = {0: "mad.wide.u32 (32x32+64)", 1: "mad.lo.u32 (32x32+32 low)", 2: "mad.lo.cc/madc.hi.cc carry chain (2 instr per MAC32)",
             3: "fma.rn.f64"}
    for kind in (0, 1, 2, 3):
        for warps in (4, 8, 16, 32):
            blocks = sm * (warps // 8 if warps >= 8 else 1)
            threads = 256 if warps >= 8 else warps * 32
            iters = 4000
            eng.probe_imad(kind, blocks, threads, 100)
            macs = [0]

            def run():
                macs[0] = eng.probe_imad(kind, blocks, threads, iters)
            ms = timed(run)
            ops = macs[0]  # one multiply-accumulate per counted op (kind 2: a lo+hi pair = one MAC32 -> counted 2x below)
            if kind == 2:
                ops //= 2  # 16 instructions per row = 8 MAC32
            res["imad"].setdefault(names[kind], {})[f"{warps}_warps_per_sm"] = round(ops / (ms * 1e-3) / 1e12, 3)
    # Fq multiplier: dependent chain per thread, many threads
    for tpsm in (256, 512, 1024, 2048):
        n = sm * tpsm
        x = torch.randint(0, 255, (n, 32), dtype=torch.uint8, device="cuda")
        x[:, 31] &= 0x1F
        y = x.flip(0).contiguous()
        iters = 2000
        eng.probe_fq_mul_dev(x, y, n, 10)
        ms = timed(lambda: eng.probe_fq_mul_dev(x, y, n, iters))
        muls = n * iters / (ms * 1e-3)
        res["fq_mul"][f"{tpsm}_threads_per_sm"] = {"gmul_per_s": round(muls / 1e9, 2), "tmac32_per_s": round(muls * 136 / 1e12, 3)}
    res["unit"] = "T ops/s (imad), see keys (fq_mul)"
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
