#!/usr/bin/env python
"""bench.py — BASELINE.json's metric for the commitment hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ipa|commit|msm|kzg|multiproof|tree] [--impl reference]

Default workload = BASELINE.json configs[1]: Pedersen/IPA commit + low_level_ipa proof at width 256, a batch
of 2^14 independent vectors per GPU; one "step" = commit all vectors, then open each at a uniform in-domain
point.  Metric: proofs/s (a proof = one committed-and-opened vector).  Under torchrun every rank runs the
same batch size (batches shard with no data-path collective: weak scaling) and `value` is the whole-job
aggregate over the max-over-ranks device time.

Synthetic data is generated ON THE DEVICE by the product itself (CRS points = k_i * G through the library's
own fixed-base kernel; scalars by rejection sampling with torch); `value` is timed with the inputs resident
in HBM, `e2e` through the C ABI's host-pointer entry points with pinned host buffers (H2D + D2H inside the
timed region).  The CPU baseline (`cpu_baseline`, and the whole `--impl reference` arm) is the oracle's
restatement of the reference's own algorithm (per-term double-and-add, utils.rs:16-19) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MAC32_PER_FQ_MUL = 136          # 8x8 + 8x8 + 8 multiply-accumulates of one Montgomery product
FQ_MUL_PER_MADD = 10            # XYZZ mixed addition: 8M + 2S
N_WIDTH = 256
WINDOW_BITS = 16
WINDOWS = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="ipa", choices=["ipa", "commit", "msm", "kzg", "multiproof", "tree"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (default: BASELINE config size)")
    ap.add_argument("--log2n", type=int, default=20, help="msm: log2 of the number of points")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--window-bits", type=int, default=0, help="fixed-base window width c of the width-256 key (default 16; up to 20)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """start of the timed region: earlier samples are kept only if the region turns out shorter than one sampling period"""
        self.first = max(0, len(self.rows) - 1)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows[getattr(self, "first", 0):] or self.rows[-1:]
        self.rows = rows
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def rand_fr_dev(torch, n, gen):
    """n uniform Fr elements (rejection sampling on the top limb; a tie with r's top limb has probability 2^-62) as
    uint8 [n, 32].  A uniform canonical value times R is uniform, so the buffer IS the Montgomery form."""
    top = R_MOD >> 192
    out = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    filled = 0
    while filled < n:
        m = int((n - filled) * 1.4) + 64
        cand = torch.randint(-(1 << 63), (1 << 63) - 1, (m, 4), dtype=torch.int64, device="cuda", generator=gen)
        cand[:, 3] &= (1 << 62) - 1
        good = cand[cand[:, 3] < top]
        k = min(len(good), n - filled)
        out[filled:filled + k] = good[:k]
        filled += k
    return out.view(torch.uint8).reshape(n, 32)


def make_points_dev(torch, eng, n, gen):
    """n random G1 points k_i * G, computed by the library's fixed-base kernel against a one-base key"""
    g = np.zeros((1, 64), dtype=np.uint8)
    g[0, :32] = np.frombuffer(((1 << 256) % P_MOD).to_bytes(32, "little"), dtype=np.uint8)        # x = 1 (Montgomery)
    g[0, 32:] = np.frombuffer(((2 << 256) % P_MOD).to_bytes(32, "little"), dtype=np.uint8)        # y = 2
    key = eng.load_key(g, window_bits=16)
    k = rand_fr_dev(torch, n, gen)
    out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    eng.commit_batch_dev(key, k, 1, n, out)
    eng.sync()
    key.free()
    return out


P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def dist_setup(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL prints a version banner on stdout at NCCL_DEBUG=VERSION; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return torch, dist, world, rank, local


def timed_steps(torch, dist, world, fn, steps, warmup, sampler=None):
    """W warm-up steps, then K steps between barrier + synchronize; CUDA events on the launching stream; max over ranks"""
    if sampler:
        sampler.start()  # nvidia-smi needs ~0.2 s to deliver its first sample: start it under the (identical) warm-up load
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.mark()   # only samples from here on count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms, clocks


def imad_peak(torch, eng):
    """measured integer-pipe peak: dependency-free mad.wide.u32 chains on every SM (tools/probe.py), T MAC32/s"""
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    eng.probe_imad(0, sm * 4, 256, 200)
    best = 0.0
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        macs = eng.probe_imad(0, sm * 4, 256, 4000)
        b.record()
        torch.cuda.synchronize()
        best = max(best, macs / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_ipa_sample(bases_h, nproofs, threads, seed=1):
    """the oracle's commit + low_level_ipa on `nproofs` vectors with `threads` host threads -> (seconds, proofs)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    rng = np.random.default_rng(seed)
    a = orc.rand_fr_buf(rng, nproofs * N_WIDTH).reshape(nproofs, N_WIDTH, 32)
    z = orc.fr_to_buf([int(v) for v in rng.integers(0, N_WIDTH, nproofs)])
    t0 = time.perf_counter()
    C = orc.commit_batch(bases_h[:N_WIDTH], a, nthreads=threads)
    orc.ipa_prove_batch(bases_h, N_WIDTH, a, C, z, nthreads=threads)
    return time.perf_counter() - t0, nproofs


def cpu_commit_sample(bases_h, n, threads, seed=1):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    rng = np.random.default_rng(seed)
    a = orc.rand_fr_buf(rng, n * N_WIDTH).reshape(n, N_WIDTH, 32)
    t0 = time.perf_counter()
    orc.commit_batch(bases_h[:N_WIDTH], a, nthreads=threads)
    return time.perf_counter() - t0, n


def cpu_msm_sample(bases_h, n, threads, seed=1):
    """reference-naive MSM (per-term double-and-add) on the first n points"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    rng = np.random.default_rng(seed)
    s = orc.rand_fr_buf(rng, n)
    t0 = time.perf_counter()
    orc.msm(bases_h[:n], s, mode="naive", nthreads=threads)
    return time.perf_counter() - t0, n


def reference_bases(n):
    """CRS for the CPU arm when no GPU produced one: the oracle's own walk of points"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    rng = np.random.default_rng(0x5EED0002)
    k0, k1 = orc.rand_fr(rng, 2)
    return orc.points_walk(k0, k1, n)


def run_reference(args):
    """--impl reference: the reference's own algorithm (oracle port) on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_cores()
    wl = args.workload
    if wl == "ipa":
        bases = reference_bases(N_WIDTH + 1)
        per_step = max(cores, 8)
        fn = lambda: cpu_ipa_sample(bases, per_step, cores)
        metric, unit, cfg = "ipa_commit_and_prove_proofs_per_s", "proofs/s", "IPA commit + low_level_ipa proof, width 256"
    elif wl == "commit":
        bases = reference_bases(N_WIDTH)
        per_step = 4 * cores
        fn = lambda: cpu_commit_sample(bases, per_step, cores)
        metric, unit, cfg = "width256_commits_per_s", "commits/s", "width-256 commit"
    elif wl == "msm":
        per_step = 256 * cores
        bases = reference_bases(per_step)
        fn = lambda: cpu_msm_sample(bases, per_step, cores)
        metric, unit, cfg = "msm_points_per_s", "points/s", f"one MSM of 2^{args.log2n} points (sample of {per_step} terms)"
    else:
        emit({"impl": "reference", "unavailable": f"no CPU arm for workload {wl}"})
        return
    for _ in range(min(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    units = 0
    for _ in range(args.steps):
        _, u = fn()
        units += u
    dt = time.perf_counter() - t0
    v = units / dt
    line = {
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integers)",
        "data": "synthetic", "config": {"workload": cfg, "sample_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                         "sample": f"{per_step} units per step x {args.steps} steps; restated reference algorithm (C++), not the arkworks binary"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def pinned(torch, shape):
    return torch.empty(shape, dtype=torch.uint8).pin_memory()


def hp(t):
    return ctypes.c_void_p(t.data_ptr())


def window_table_bytes(c, nbases):
    return nbases * ((256 + c - 1) // c) * (1 << (c - 1)) * 64


def pick_window_bits(torch, nbases):
    """largest fixed-base window width whose tables fit this GPU's FREE memory with room for the batch (the tables are the
    memory-for-work knob of the design: 13 table additions per scalar at c = 20 / 112 GB, 16 at c = 16 / 8.6 GB)"""
    need = window_table_bytes(20, nbases) + (12 << 30)
    free, total = torch.cuda.mem_get_info()
    # back-to-back runs: the previous process's memory may still be on its way back to the driver — wait for it (bounded)
    # instead of silently benchmarking a smaller table
    deadline = time.perf_counter() + 20.0
    while free < need <= total and time.perf_counter() < deadline:
        time.sleep(0.5)
        free, total = torch.cuda.mem_get_info()
    for c in (20, 19, 18, 16):
        if window_table_bytes(c, nbases) + (12 << 30) <= free:
            return c
    return 16


def run_native(args):
    global WINDOW_BITS, WINDOWS
    torch, dist, world, rank, local = dist_setup(args)
    WINDOW_BITS = args.window_bits or pick_window_bits(torch, N_WIDTH + 1)
    WINDOWS = (256 + WINDOW_BITS - 1) // WINDOW_BITS
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: verkle_kzg_b200 has no CPU path")
    from verkle_kzg_b200 import Engine, _lib
    from verkle_kzg_b200._lib import check
    eng = Engine(local, stream=torch.cuda.current_stream().cuda_stream)
    L = _lib.lib()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x5EED0002)  # same CRS on every rank
    wl = args.workload
    sampler = ClockSampler(local) if rank == 0 else None
    peak = imad_peak(torch, eng)
    extra = {}

    if wl in ("ipa", "commit", "kzg"):
        B = args.batch or (1 << 14)
        bases = make_points_dev(torch, eng, N_WIDTH + 1, gen)
        t_key = time.perf_counter()
        key = eng.load_key_dev(bases, N_WIDTH, d_q=bases[N_WIDTH:] if wl == "ipa" else None, window_bits=WINDOW_BITS)
        extra["key_load_s"] = round(time.perf_counter() - t_key, 2)
        gen.manual_seed(0x5EED1000 + rank)
        a = rand_fr_dev(torch, B * N_WIDTH, gen).reshape(B, N_WIDTH, 32)
        zi = torch.randint(0, N_WIDTH, (B,), device="cuda", generator=gen)
        # F::from(z) in Montgomery form: z * R mod r, exact on the host for the 256 possible values
        tab = np.stack([np.frombuffer(((v << 256) % R_MOD).to_bytes(32, "little"), dtype=np.uint8) for v in range(N_WIDTH)])
        z = torch.from_numpy(tab).cuda()[zi].contiguous()
        C = torch.empty((B, 64), dtype=torch.uint8, device="cuda")
        Lr = torch.empty((B, 8, 64), dtype=torch.uint8, device="cuda")
        Rr = torch.empty((B, 8, 64), dtype=torch.uint8, device="cuda")
        tip = torch.empty((B, 32), dtype=torch.uint8, device="cuda")
        y = torch.empty((B, 32), dtype=torch.uint8, device="cuda")
        pf = torch.empty((B, 64), dtype=torch.uint8, device="cuda")
        a_h, z_h = pinned(torch, (B, N_WIDTH, 32)), pinned(torch, (B, 32))
        a_h.copy_(a)
        z_h.copy_(z)
        C_h, L_h, R_h = pinned(torch, (B, 64)), pinned(torch, (B, 8, 64)), pinned(torch, (B, 8, 64))
        tip_h, y_h = pinned(torch, (B, 32)), pinned(torch, (B, 32))
        kid = ctypes.c_uint32(key.id)
        if wl == "ipa":
            def step():
                eng.commit_batch_dev(key, a, N_WIDTH, B, C)
                eng.ipa_prove_batch_dev(key, a, z, C, B, Lr, Rr, tip, y)

            def step_e2e():
                # host buffers in, host buffers out: commit + open through the batch entry point (one upload of the rows)
                check(L.vkzg_ipa_commit_prove_batch(eng._ctx, kid, hp(a_h), hp(z_h), ctypes.c_uint64(B), hp(C_h), hp(L_h), hp(R_h),
                                                    hp(tip_h), hp(y_h)), "commit_prove")
            madds_per_unit = (N_WIDTH + 8 * 2 * (N_WIDTH // 2 + 1)) * WINDOWS          # commit + 8 rounds of two 129-term MSMs
            launches_timed = 9
            h2d = a_h.numel() + z_h.numel()
            d2h = C_h.numel() + L_h.numel() + R_h.numel() + tip_h.numel() + y_h.numel()
            metric, unit = "ipa_commit_and_prove_proofs_per_s", "proofs/s"
            cfg = {"workload": "configs[1]: Pedersen/IPA commit + low_level_ipa proof, width 256, batch 2^14 vectors per GPU",
                   "batch_per_gpu": B, "width": N_WIDTH, "points": "uniform in-domain index per vector"}
            cpu_fn = cpu_ipa_sample
        elif wl == "commit":
            def step():
                eng.commit_batch_dev(key, a, N_WIDTH, B, C)

            def step_e2e():
                check(L.vkzg_commit_batch(eng._ctx, kid, hp(a_h), ctypes.c_uint32(N_WIDTH), ctypes.c_uint64(B), hp(C_h)), "commit")
            madds_per_unit = N_WIDTH * WINDOWS
            launches_timed = 1
            h2d, d2h = a_h.numel(), C_h.numel()
            metric, unit = "width256_commits_per_s", "commits/s"
            cfg = {"workload": "width-256 Pedersen/KZG commit (M1), batch 2^14 vectors per GPU", "batch_per_gpu": B, "width": N_WIDTH}
            cpu_fn = cpu_commit_sample
        else:
            def step():
                eng.commit_batch_dev(key, a, N_WIDTH, B, C)
                eng.kzg_open_batch_dev(key, a, N_WIDTH, z, B, pf, y)

            def step_e2e():   # commit + open of the same vectors: one upload of the rows
                check(L.vkzg_kzg_commit_open_batch(eng._ctx, kid, hp(a_h), ctypes.c_uint32(N_WIDTH), ctypes.c_uint32(0), hp(z_h),
                                                   ctypes.c_uint64(B), hp(C_h), hp(L_h), hp(y_h)), "commit+open")
            madds_per_unit = 2 * N_WIDTH * WINDOWS
            launches_timed = 2
            h2d, d2h = a_h.numel() + z_h.numel(), 2 * C_h.numel() + y_h.numel()
            metric, unit = "kzg_commit_and_open_per_s", "openings/s"
            cfg = {"workload": "configs[0] at batch: KZG commit + single-point open, width 256, batch 2^14 per GPU", "batch_per_gpu": B}
            cpu_fn = None
        units_per_step = B
        bases_h = bases.cpu().numpy()
    elif wl == "msm":
        n = 1 << args.log2n
        per = n // world                                                        # point-range sharding (configs[3])
        gen.manual_seed(0x5EED0004 + rank)
        bases = make_points_dev(torch, eng, per, gen)
        t_key = time.perf_counter()
        key = eng.load_key_dev(bases, per, kind=_lib.KEY_MSM)
        extra["key_load_s"] = round(time.perf_counter() - t_key, 2)
        extra["msm_table_gb"] = round(key.table_bytes / 1e9, 2)
        s = rand_fr_dev(torch, per, gen)
        part = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
        allp = torch.empty((world, 64), dtype=torch.uint8, device="cuda")
        out = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
        s_h, out_h = pinned(torch, (per, 32)), pinned(torch, (1, 64))
        s_h.copy_(s)
        kid = ctypes.c_uint32(key.id)

        def step():
            eng.msm_dev(key, s, per, part)
            if world > 1:                                                       # partial sums combined over NCCL
                dist.all_gather_into_tensor(allp, part)
                eng.g1_sum_dev(allp, world, out)

        def step_e2e():
            check(L.vkzg_msm(eng._ctx, kid, hp(s_h), ctypes.c_uint64(per), hp(out_h)), "msm")
            if world > 1:
                part.copy_(out_h, non_blocking=True)
                dist.all_gather_into_tensor(allp, part)
                eng.g1_sum_dev(allp, world, out)
                out_h.copy_(out)
        madds_per_unit = 16                                                      # MSM keys: c = 16 -> 16 signed digits per scalar
        launches_timed = 1
        units_per_step = per
        h2d, d2h = s_h.numel(), 64
        metric, unit = "msm_points_per_s", "points/s"
        cfg = {"workload": f"configs[3]: one KZG-commit MSM of 2^{args.log2n} points, point-range sharded over {world} GPU(s)",
               "points_per_gpu": per}
        cpu_fn = cpu_msm_sample
        bases_h = None
    elif wl == "multiproof":
        m = args.batch or (1 << 12)                                             # configs[2]: 2^12 openings aggregated
        bases = make_points_dev(torch, eng, N_WIDTH + 1, gen)
        key = eng.load_key_dev(bases, N_WIDTH, d_q=bases[N_WIDTH:], window_bits=WINDOW_BITS)
        gen.manual_seed(0x5EED3000 + rank)
        f = rand_fr_dev(torch, m * N_WIDTH, gen).reshape(m, N_WIDTH, 32)
        zi = torch.randint(0, N_WIDTH, (m,), device="cuda", generator=gen)
        Cq = torch.empty((m, 64), dtype=torch.uint8, device="cuda")
        eng.commit_batch_dev(key, f, N_WIDTH, m, Cq)
        yq = f[torch.arange(m, device="cuda"), zi].contiguous()
        f_h = pinned(torch, (m, N_WIDTH, 32))
        f_h.copy_(f)
        C_h, y_h = Cq.cpu().contiguous(), yq.cpu().contiguous()
        z_h = zi.cpu().numpy().astype(np.uint64)
        D_h, L_h, R_h = pinned(torch, (64,)), pinned(torch, (8, 64)), pinned(torch, (8, 64))
        tip_h, yo_h = pinned(torch, (32,)), pinned(torch, (32,))
        kid = ctypes.c_uint32(key.id)
        zp = z_h.ctypes.data_as(ctypes.c_void_p)

        def step():
            check(L.vkzg_multiproof_prove_dev(eng._ctx, kid, ctypes.c_int32(0), ctypes.c_void_p(f.data_ptr()), hp(C_h), zp, hp(y_h),
                                              ctypes.c_uint64(m), hp(D_h), hp(L_h), hp(R_h), hp(tip_h), hp(yo_h)), "multiproof_dev")

        def step_e2e():
            check(L.vkzg_multiproof_prove(eng._ctx, kid, ctypes.c_int32(0), hp(f_h), hp(C_h), zp, hp(y_h), ctypes.c_uint64(m), hp(D_h),
                                          hp(L_h), hp(R_h), hp(tip_h), hp(yo_h)), "multiproof")
        madds_per_unit = (2 * N_WIDTH + 8 * 2 * (N_WIDTH // 2 + 1)) * WINDOWS   # D, E commits + the final IPA opening
        launches_timed = 10
        units_per_step = 1
        h2d, d2h = f_h.numel() + m * (64 + 8 + 32), 64 + 2 * 8 * 64 + 64
        metric, unit = "ipa_multiproofs_per_s", "multiproofs/s"
        cfg = {"workload": f"configs[2]: IPA multiproof aggregating {m} openings at width 256 (one multiproof per step per GPU; replicas across GPUs)",
               "openings_per_multiproof": m}
        cpu_fn = None
        bases_h = None
    elif wl == "tree":
        nkeys = args.batch or (1 << 20)                                         # configs[4]: bulk insert of 2^20 keys
        from verkle_kzg_b200.tree import build_levels
        from verkle_kzg_b200.sharding import split_range
        bases = make_points_dev(torch, eng, N_WIDTH, gen)
        key = eng.load_key_dev(bases, N_WIDTH, window_bits=WINDOW_BITS)
        rng = np.random.default_rng(0x5EED0005)
        keys_all = rng.integers(0, 256, (nkeys, 32), dtype=np.uint8)
        vals_all = rng.integers(0, 256, (nkeys, 32), dtype=np.uint8)
        lo, cnt = split_range(256, world, rank)                                 # shard by top-level child index (whole subtrees per rank)
        sel = (keys_all[:, 0] >= lo) & (keys_all[:, 0] < lo + cnt)
        t0 = time.perf_counter()
        levels = build_levels(keys_all[sel], vals_all[sel], 256)                # host pointer structure -> level lists (not timed: out of the path)
        host_build_s = time.perf_counter() - t0
        counts = [len(lv["row_ptr"]) - 1 for lv in levels]
        total_nodes = sum(counts)
        dev = [dict(rp=torch.from_numpy(lv["row_ptr"].astype(np.int32)).cuda(), sl=torch.from_numpy(lv["slot"].astype(np.int16)).cuda(),
                    ch=torch.from_numpy(lv["child"]).cuda(), li=torch.from_numpy(np.ascontiguousarray(lv["lit"])).cuda(),
                    n=c, t=len(lv["slot"])) for lv, c in zip(levels, counts)]
        nodes = torch.empty((total_nodes, 64), dtype=torch.uint8, device="cuda")
        allp = torch.empty((world, 64), dtype=torch.uint8, device="cuda")
        root = torch.empty((1, 64), dtype=torch.uint8, device="cuda")

        def step():
            off = 0
            for d in dev:
                eng.tree_level_dev(key, d["rp"], d["n"], d["sl"] if d["t"] else None, d["ch"] if d["t"] else None,
                                   d["li"] if d["t"] else None, d["t"], nodes, nodes[off:])
                off += d["n"]
            if world > 1:   # one exchange at the root: per-rank partial roots (disjoint child slots) are added
                dist.all_gather_into_tensor(allp, nodes[total_nodes - 1:total_nodes])
                eng.g1_sum_dev(allp, world, root)

        from verkle_kzg_b200.tree import NativeVerkleTree
        kk, vv = keys_all[sel], vals_all[sel]

        def step_e2e():
            # configs[4] end to end: bulk insert into the native host tree (libvkzg, Node::insert semantics) + recommit to the root
            t = NativeVerkleTree(32, 256)
            t.insert_many(kk, vv)
            t.commitment(eng, key)
            t.close()
        terms = sum(len(lv["slot"]) for lv in levels)
        madds_per_unit = terms * WINDOWS / max(1, int(sel.sum()))               # upper bound: zero digits are skipped at run time
        launches_timed = len(levels)
        units_per_step = int(sel.sum())
        # e2e traffic of vkzg_tree_commit: extensions travel compact (32-byte stem + unit + 32-byte value; the device writes
        # their CSR rows), internal levels as row_ptr + (slot, child) per term; every node commitment is cached back
        h2d = counts[1] * 65 + sum(4 * (c + 1) + 6 * len(lv["slot"]) for lv, c in zip(levels[2:], counts[2:]))
        d2h = 64 * sum(counts[1:])
        metric, unit = "verkle_tree_commit_keys_per_s", "keys/s"
        cfg = {"workload": f"configs[4]: verkle tree of {nkeys} random 32-byte keys, every node recommitted level by level up to the root",
               "keys_this_rank": units_per_step, "nodes_per_level": counts, "terms": terms, "host_flatten_seconds_not_timed": round(host_build_s, 1)}
        cpu_fn = None
        bases_h = None
    else:
        raise SystemExit(f"unknown workload {wl}")

    cfg.update(extra)
    if wl != "msm":
        cfg["window_bits"] = WINDOW_BITS
        cfg["window_table_gb"] = round(key.table_bytes / 1e9, 1)
    cfg["l2_policy"] = "tables (GBs) and inputs (>= 128 MB) exceed the 126 MB L2; no flush needed"
    # ---- device-resident timing (the `value`): W warm-up steps, K timed steps
    l0 = eng.launches
    ms, clocks = timed_steps(torch, dist, world, step, args.steps, args.warmup, sampler)
    launches = (eng.launches - l0) * args.steps // (args.steps + args.warmup)
    # ---- the dominant kernel's own launch durations (roofline): every launch bracketed by a CUDA event pair on its stream.
    #      Taken in a second pass of K steps with the IPA half-batches on ONE stream, so that a bracket times one kernel alone
    #      (in the timed region above the two half-batches interleave on two streams and brackets would overlap).
    eng.set_option(eng.OPT_IPA_TWO_STREAMS, 0)
    step()
    eng.kernel_timing(True)
    ms_k, _ = timed_steps(torch, dist, world, step, args.steps, 0)
    kn, kms = eng.kernel_timing_read()
    eng.kernel_timing(False)
    eng.set_option(eng.OPT_IPA_TWO_STREAMS, 1)
    if world > 1:
        tu = torch.tensor([units_per_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tu)
        units_all = float(tu.item())
    else:
        units_all = float(units_per_step)
    total_units = units_all * args.steps
    value = total_units / (ms * 1e-3)
    # ---- end to end through the host-pointer C ABI
    # (2 warm-up steps: the host-pointer path has its own scratch and staging buffers to put into the stream-ordered pool; two
    #  timed repetitions of K/2 steps, the faster one reported: a one-off pool growth or a neighbour's PCIe burst inside a
    #  2-step window otherwise shows up as a 30 % outlier — seen once in ~20 runs)
    e2e_steps = max(1, args.steps // 2)
    ms_e2e = min(timed_steps(torch, dist, world, step_e2e, e2e_steps, 2)[0], timed_steps(torch, dist, world, step_e2e, e2e_steps, 0)[0])
    e2e_value = units_all * e2e_steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    macs_per_launch_set = units_per_step * madds_per_unit * FQ_MUL_PER_MADD * MAC32_PER_FQ_MUL
    achieved = macs_per_launch_set * args.steps / (kms * 1e-3) / 1e12 if kms > 0 else None
    # The same kernel against the HBM roofline (algorithmic bytes = one 64-byte table point per addition): the gather
    # stream is a few per cent of the measured copy bandwidth, i.e. the kernel is not memory-bound.
    hbm_peak, hbm_src = measured_hbm_peak()
    hbm_view = None
    if kn and kms:
        gbs = (units_per_step * madds_per_unit * args.steps) * 64.0 / (kms * 1e-3) / 1e9
        hbm_view = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src}
    # SURVEY.md section 8d's per-unit figures (8-bit windows, Jacobian 11-mul additions): the work the REFERENCE algorithm
    # would need, not what this kernel executes (wider windows need fewer additions), so its "frac" can exceed 1.
    survey_mac = {"ipa": 12.3e6 + 98.9e6, "commit": 12.3e6, "kzg": 12.3e6 + 0.17e6,
                  "msm": {16: 35.4e3, 18: 31.0e3, 20: 26.1e3}.get(getattr(args, "log2n", 20))}.get(wl)
    survey_acct = None
    if survey_mac and kms:
        a = units_per_step * survey_mac * args.steps / (kms * 1e-3) / 1e12
        survey_acct = {"mac32_per_unit": survey_mac, "achieved": a, "frac": a / peak if peak else None,
                       "note": "reference-algorithm work per unit; > 1 means fewer operations were executed than that algorithm needs"}
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if wl in ("msm", "tree") else "weak", "vs_baseline": None,
        "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic", "config": cfg,
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "timing": f"faster of 2 repetitions of {e2e_steps} steps after 2 warm-up steps, CUDA events, max over ranks"},
        "roofline": {
            "bound": "int32",
            "bound_note": "integer multiply pipe (IMAD): 254-bit modular arithmetic is neither hbm- nor tensor-bound; the hbm view of the same kernel is under hbm_view",
            "kernel": "k_msm_bucket" if wl == "msm" else "k_fixed_base_msm",
            "achieved": achieved, "peak": peak, "unit": "TMAC32/s", "frac": (achieved / peak) if achieved else None,
            "peak_source": "measured live: dependency-free mad.wide.u32 chains on all SMs (vkzg_probe_imad_dev); MEASURED_PEAKS.json has no integer figure",
            "work_model": f"{madds_per_unit} mixed additions per unit x {FQ_MUL_PER_MADD} Fq-mul x {MAC32_PER_FQ_MUL} MAC32",
            "survey_accounting": survey_acct,
            "kernel_launches_timed": kn, "kernel_ms_total": kms, "kernel_share_of_step": kms / ms_k if ms_k else None,
            "one_stream_ms_per_step": ms_k / args.steps,
            # DRAM bytes per launch: 128 B per table addition as ncu measured it on this kernel (dram__bytes_read.sum +
            # dram__bytes_write.sum of one --set full capture, profiles/r01_ncu_full_summary_final.json: 1.785 GB for a
            # launch of 13.7 M additions; every 64-byte point is fetched at 128-byte granularity) x additions per launch
            "traffic": (units_per_step * madds_per_unit * args.steps / kn) * 128.0 if kn else None,
            "algorithmic_bytes_per_launch": (units_per_step * madds_per_unit * args.steps / kn) * 64.0 if kn else None,
            "hbm_view": hbm_view,
            "ncu": "sm__pipe_fmaheavy_cycles_active 85 % (k_fixed_base_msm), 86 % (k_msm_bucket); DRAM read ~10 % of peak (profiles/)",
        },
    }
    if cpu_fn and not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
        cores = cpu_cores()
        if wl == "msm":
            nsmp = 128 * cores
            bh = make_points_dev(torch, eng, nsmp, gen).cpu().numpy()
            dt, u = cpu_fn(bh, nsmp, cores)
            sample = f"{nsmp} of the 2^{args.log2n} terms (naive per-term double-and-add scales linearly)"
        else:
            nsmp = 4 * cores if wl == "ipa" else 16 * cores
            dt, u = cpu_fn(bases_h, nsmp, cores)
            sample = f"{nsmp} of the {units_per_step} vectors"
        line["cpu_baseline"] = {"value": u / dt, "unit": unit, "cores": cores, "kind": "port",
                                "sample": sample + "; restated reference algorithm (C++ oracle), not the arkworks binary"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def measured_hbm_peak():
    """HBM copy bandwidth from the driver-written MEASURED_PEAKS.json, else B200_PROFILING.md's fallback."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner, torchrun
    notices) was diverted to stderr for the lifetime of the process"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
