#!/usr/bin/env python
"""bench.py — BASELINE.json's metrics for the commitment hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ipa|commit|msm|kzg|multiproof|tree] [--impl reference]
                    [--no-also] [--no-check]

Headline (default workload) = BASELINE.json configs[1]: Pedersen/IPA commit + low_level_ipa proof at width 256, a
batch of 2^14 independent vectors per GPU; one "step" = commit all vectors, then open each at a uniform in-domain
point.  Metric: proofs/s (a proof = one committed-and-opened vector).  Under torchrun every rank runs the same batch
size (batches shard with no data-path collective: weak scaling) and `value` is the whole-job aggregate over the
max-over-ranks device time.

The default run also measures BASELINE's other named metrics and puts them into the same JSON line under `also`
(msm_2p20 — under torchrun point-range sharded with an all_gather of the partial sums —, commit_w256, kzg_open,
multiproof_2p12, tree_2p20, and the headline again at the library's default window width c = 16), each with
`value`, `e2e`, `roofline.frac`, `cpu_baseline` and `checked`.

Every workload CHECKS the outputs of its timed region before its numbers are reported (`checked`): all proofs of the
last step go through the device verifier, samples are compared byte for byte with the oracle; a mismatch exits
non-zero.  The oracle is used only there and in the CPU-baseline legs.

Synthetic data is generated ON THE DEVICE by the product itself (CRS points = k_i * G through the library's own
fixed-base kernel; scalars by rejection sampling with torch); `value` is timed with the inputs resident in HBM, `e2e`
through the C ABI's host-pointer entry points with pinned host buffers (H2D + D2H inside the timed region).  The CPU
baseline (`cpu_baseline`, and the whole `--impl reference` arm) is the oracle's restatement of the reference's own
algorithm (per-term double-and-add, utils.rs:16-19) on the host cores.
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
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
MAC32_PER_FQ_MUL = 136          # 8x8 + 8x8 + 8 multiply-accumulates of one Montgomery product
FQ_MUL_PER_MADD = 10            # XYZZ mixed addition: 8M + 2S (the work model of DESIGN.md section 3, whatever the kernel executes)
N_WIDTH = 256
WORKLOADS = ["ipa", "commit", "msm", "kzg", "multiproof", "mpbatch", "tree"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="ipa", choices=WORKLOADS)
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (default: BASELINE config size)")
    ap.add_argument("--log2n", type=int, default=20, help="msm: log2 of the number of points")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="headline workload only (no `also` block)")
    ap.add_argument("--no-check", action="store_true", help="skip the output checks (profiling runs)")
    ap.add_argument("--window-bits", type=int, default=0, help="fixed-base window width c of the width-256 key (default: largest that fits, up to 20)")
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


def dist_setup(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # (NCCL_DEBUG is left as the caller set it: stdout is diverted to stderr for the lifetime of the process, so NCCL's
        #  banner cannot end up next to the one JSON line — see emit())
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


def _orc():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    return orc


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _pool_map(fn, items, threads):
    """the oracle's C entry points release the GIL (ctypes): independent units run on `threads` host threads"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(fn, items))


def cpu_ipa_sample(bases_h, nproofs, threads, seed=1):
    """the oracle's commit + low_level_ipa on `nproofs` vectors with `threads` host threads -> (seconds, proofs)"""
    orc = _orc()
    rng = np.random.default_rng(seed)
    a = orc.rand_fr_buf(rng, nproofs * N_WIDTH).reshape(nproofs, N_WIDTH, 32)
    z = orc.fr_to_buf([int(v) for v in rng.integers(0, N_WIDTH, nproofs)])
    t0 = time.perf_counter()
    C = orc.commit_batch(bases_h[:N_WIDTH], a, nthreads=threads)
    orc.ipa_prove_batch(bases_h, N_WIDTH, a, C, z, nthreads=threads)
    return time.perf_counter() - t0, nproofs


def cpu_commit_sample(bases_h, n, threads, seed=1):
    orc = _orc()
    rng = np.random.default_rng(seed)
    a = orc.rand_fr_buf(rng, n * N_WIDTH).reshape(n, N_WIDTH, 32)
    t0 = time.perf_counter()
    orc.commit_batch(bases_h[:N_WIDTH], a, nthreads=threads)
    return time.perf_counter() - t0, n


def cpu_msm_sample(bases_h, n, threads, seed=1):
    """reference-naive MSM (per-term double-and-add) on the first n points"""
    orc = _orc()
    rng = np.random.default_rng(seed)
    s = orc.rand_fr_buf(rng, n)
    t0 = time.perf_counter()
    orc.msm(bases_h[:n], s, mode="naive", nthreads=threads)
    return time.perf_counter() - t0, n


def cpu_kzg_sample(bases_h, n, threads, seed=1):
    """KZG::commit + KZG::prove_point per vector (benches/kzg.rs:61-75), vectors spread over the host threads"""
    orc = _orc()
    rng = np.random.default_rng(seed)
    a = orc.rand_fr_buf(rng, n * N_WIDTH).reshape(n, N_WIDTH, 32)
    z = orc.fr_to_buf([int(v) for v in rng.integers(0, N_WIDTH, n)])
    t0 = time.perf_counter()
    orc.commit_batch(bases_h[:N_WIDTH], a, nthreads=threads)
    _pool_map(lambda i: orc.kzg_prove(bases_h[:N_WIDTH], a[i], z[i]), range(n), threads)
    return time.perf_counter() - t0, n


def cpu_multiproof_sample(bases_h, m, threads, seed=1, reps=1):
    """prove_multiproof over m openings (multiproof.rs:99-176, benches/ipa.rs:111-132); the query commitments are inputs and
    are not timed.  `reps` independent multiproofs run side by side on the host threads."""
    orc = _orc()
    rng = np.random.default_rng(seed)
    f = orc.rand_fr_buf(rng, m * N_WIDTH).reshape(m, N_WIDTH, 32)
    z = rng.integers(0, N_WIDTH, m).astype(np.uint64)
    y = f[np.arange(m), z.astype(np.int64)]
    C = np.tile(bases_h[:1], (m, 1))   # the transcript only hashes C: any valid points cost the same
    t0 = time.perf_counter()
    _pool_map(lambda _i: orc.multiproof_prove("ipa", bases_h, N_WIDTH, f, C, z, y), range(reps), min(threads, reps))
    return time.perf_counter() - t0, reps


def cpu_tree_sample(bases_h, nkeys, threads, seed=1):
    """VerkleTree insert + commitment (verkle-tree/src/lib.rs:106-137) on independent trees of nkeys keys, one per thread"""
    orc = _orc()
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 256, (threads, nkeys, 32), dtype=np.uint8)
    vals = rng.integers(0, 256, (threads, nkeys, 32), dtype=np.uint8)
    t0 = time.perf_counter()
    _pool_map(lambda i: orc.tree_commit(bases_h[:N_WIDTH], keys[i], vals[i]), range(threads), threads)
    return time.perf_counter() - t0, nkeys * threads


def reference_bases(n):
    """CRS for the CPU arm when no GPU produced one: the oracle's own walk of points"""
    orc = _orc()
    rng = np.random.default_rng(0x5EED0002)
    k0, k1 = orc.rand_fr(rng, 2)
    return orc.points_walk(k0, k1, n)


CPU_ARMS = {
    # workload: (sample function, units per step as a function of cores, metric, unit, description)
    "ipa": (cpu_ipa_sample, lambda c: max(c, 8), "ipa_commit_and_prove_proofs_per_s", "proofs/s", "IPA commit + low_level_ipa proof, width 256"),
    "commit": (cpu_commit_sample, lambda c: 4 * c, "width256_commits_per_s", "commits/s", "width-256 commit"),
    "msm": (cpu_msm_sample, lambda c: 256 * c, "msm_points_per_s", "points/s", "one MSM (sample of the terms; the reference's per-term double-and-add is linear in n)"),
    "kzg": (cpu_kzg_sample, lambda c: 4 * c, "kzg_commit_and_open_per_s", "openings/s", "KZG commit + single-point open, width 256"),
    "multiproof": (None, None, "ipa_multiproofs_per_s", "multiproofs/s", "IPA multiproof over 2^12 openings, width 256"),
    "mpbatch": (None, None, "ipa_multiproofs_per_s", "multiproofs/s", "IPA multiproof over 2^12 openings, width 256"),
    "tree": (cpu_tree_sample, lambda c: 256, "verkle_tree_commit_keys_per_s", "keys/s", "verkle tree insert + commitment (independent 256-key trees, one per thread)"),
}


def run_reference(args):
    """--impl reference: the reference's own algorithm (oracle port) on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_cores()
    wl = args.workload
    fn_s, per_fn, metric, unit, cfg = CPU_ARMS[wl]
    if wl in ("multiproof", "mpbatch"):
        m = (args.batch or (1 << 12)) if wl == "multiproof" else 1 << 12
        bases = reference_bases(N_WIDTH + 1)
        reps = min(cores, 4)
        per_step = reps
        fn = lambda: cpu_multiproof_sample(bases, m, cores, reps=reps)
    elif wl == "tree":
        per_step = 256 * cores
        bases = reference_bases(N_WIDTH + 1)
        fn = lambda: cpu_tree_sample(bases, 256, cores)
    else:
        per_step = per_fn(cores)
        bases = reference_bases(per_step if wl == "msm" else N_WIDTH + 1)
        fn = lambda: fn_s(bases, per_step, cores)
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


class CheckFailed(Exception):
    pass


def _need(cond, what):
    if not cond:
        raise CheckFailed(what)


class Ctx:
    """what every workload shares: the process group, the library context, the CRS and its (cached) window key"""

    def __init__(self, args):
        self.args = args
        self.torch, self.dist, self.world, self.rank, self.local = dist_setup(args)
        torch = self.torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: verkle_kzg_b200 has no CPU path")
        from verkle_kzg_b200 import Engine, _lib
        self._lib = _lib
        self.eng = Engine(self.local, stream=torch.cuda.current_stream().cuda_stream)
        self.L = _lib.lib()
        self.gen = torch.Generator(device="cuda")
        self.gen.manual_seed(0x5EED0002)  # same CRS on every rank
        self.peak = imad_peak(torch, self.eng)
        self.bases = make_points_dev(torch, self.eng, N_WIDTH + 1, self.gen)       # 256 bases + Q
        self.bases_h = self.bases.cpu().numpy()
        self.window_bits = args.window_bits or pick_window_bits(torch, N_WIDTH + 1)
        self._key, self._key_c, self.key_load_s = None, 0, None

    def key257(self, c):
        if self._key is not None and self._key_c != c:
            self._key.free()
            self._key = None
        if self._key is None:
            t = time.perf_counter()
            self._key = self.eng.load_key_dev(self.bases, N_WIDTH, d_q=self.bases[N_WIDTH:], window_bits=c)
            self.eng.sync()
            self.key_load_s = round(time.perf_counter() - t, 2)
            self._key_c = c
        return self._key

    def drop_key(self):
        if self._key is not None:
            self._key.free()
            self._key = None
        self.torch.cuda.empty_cache()


def build_workload(ctx, wl, c):
    """-> dict describing one workload: step / step_e2e closures, the work model, the CPU sample and the output check"""
    torch, dist, world, rank, eng, L, gen, args = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.eng, ctx.L, ctx.gen, ctx.args
    check = ctx._lib.check
    W = (256 + c - 1) // c
    cores = cpu_cores()
    w = {"wl": wl, "extra": {}, "kernel": "k_fixed_base_msm", "scaling": "weak", "cleanup": lambda: None}

    if wl in ("ipa", "commit", "kzg"):
        B = args.batch or (1 << 14)
        key = ctx.key257(c)
        w["extra"]["key_load_s"] = ctx.key_load_s
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
        samples = sorted({0, B // 2 - 1, B // 2, B - 1} & set(range(B)))     # both half-batches of the two-stream split
        if wl == "ipa":
            def step():
                eng.commit_batch_dev(key, a, N_WIDTH, B, C)
                eng.ipa_prove_batch_dev(key, a, z, C, B, Lr, Rr, tip, y)

            def step_e2e():
                # host buffers in, host buffers out: commit + open through the batch entry point (one upload of the rows)
                check(L.vkzg_ipa_commit_prove_batch(eng._ctx, kid, hp(a_h), hp(z_h), ctypes.c_uint64(B), hp(C_h), hp(L_h), hp(R_h),
                                                    hp(tip_h), hp(y_h)), "commit_prove")

            def checker():
                """every proof of the last timed step through the device verifier (ipa/mod.rs:404-421: prove -> verify), four of
                them against the oracle byte for byte, and the e2e path's outputs equal to the device-resident path's"""
                orc = _orc()
                Cn, Ln, Rn, tn, yn = C.cpu().numpy(), Lr.cpu().numpy(), Rr.cpu().numpy(), tip.cpu().numpy(), y.cpu().numpy()
                zn, an = z_h.numpy(), a_h.numpy()
                eng.ipa_verify_batch(key, zn[:256], Cn[:256], Ln[:256], Rn[:256], tn[:256], yn[:256])   # warm-up of the verifier's kernels
                t_v = time.perf_counter()
                ok = eng.ipa_verify_batch(key, zn, Cn, Ln, Rn, tn, yn)
                t_v = time.perf_counter() - t_v
                _need(bool(ok.all()), f"{int((~ok).sum())} of {B} proofs of the timed region fail the device verifier")
                for i in samples:
                    eC = orc.commit_batch(ctx.bases_h[:N_WIDTH], an[i:i + 1])[0]
                    eL, eR, etip, ey = orc.ipa_prove(ctx.bases_h, N_WIDTH, an[i], eC, zn[i])
                    _need((Cn[i] == eC).all() and (Ln[i] == eL).all() and (Rn[i] == eR).all() and (tn[i] == etip).all()
                          and (yn[i] == ey).all(), f"proof {i} differs from the oracle")
                _need((C_h.numpy() == Cn).all() and (L_h.numpy() == Ln).all() and (R_h.numpy() == Rn).all()
                      and (tip_h.numpy() == tn).all() and (y_h.numpy() == yn).all(), "e2e outputs differ from the device-resident outputs")
                # (verifier rate: host buffers through vkzg_ipa_verify_batch, wall clock of one call — SURVEY 8f-3, reported, not a bench metric)
                return {"verified": B, "oracle_samples": len(samples), "e2e_equals_device": True, "verify_proofs_per_s_e2e": round(B / t_v, 1)}
            madds = (N_WIDTH + 8 * 2 * (N_WIDTH // 2 + 1)) * W          # commit + 8 rounds of two 129-term MSMs
            h2d = a_h.numel() + z_h.numel()
            d2h = C_h.numel() + L_h.numel() + R_h.numel() + tip_h.numel() + y_h.numel()
            w.update(metric="ipa_commit_and_prove_proofs_per_s", unit="proofs/s",
                     cfg={"workload": "configs[1]: Pedersen/IPA commit + low_level_ipa proof, width 256, batch 2^14 vectors per GPU",
                          "batch_per_gpu": B, "width": N_WIDTH, "points": "uniform in-domain index per vector"},
                     cpu=lambda: cpu_ipa_sample(ctx.bases_h, 4 * cores, cores) + (f"{4 * cores} of the {B} vectors",))
        elif wl == "commit":
            def step():
                eng.commit_batch_dev(key, a, N_WIDTH, B, C)

            def step_e2e():
                check(L.vkzg_commit_batch(eng._ctx, kid, hp(a_h), ctypes.c_uint32(N_WIDTH), ctypes.c_uint64(B), hp(C_h)), "commit")

            def checker():
                orc = _orc()
                Cn, an = C.cpu().numpy(), a_h.numpy()
                for i in samples:
                    _need((Cn[i] == orc.commit_batch(ctx.bases_h[:N_WIDTH], an[i:i + 1])[0]).all(), f"commit {i} differs from the oracle")
                _need((C_h.numpy() == Cn).all(), "e2e outputs differ from the device-resident outputs")
                # size-independent: commit is linear — commit(a_0 + a_half) = commit(a_0) + commit(a_half)
                both = eng.fr_vector_op("add", an[0], an[B // 2]).reshape(1, N_WIDTH, 32)
                _need((eng.commit_batch(key, both)[0] == orc.g1_add(Cn[0], Cn[B // 2])).all(), "commit(a0 + a_half) != commit(a0) + commit(a_half)")
                return {"oracle_samples": len(samples), "linearity": True, "e2e_equals_device": True}
            madds = N_WIDTH * W
            h2d, d2h = a_h.numel(), C_h.numel()
            w.update(metric="width256_commits_per_s", unit="commits/s",
                     cfg={"workload": "width-256 Pedersen/KZG commit (M1), batch 2^14 vectors per GPU", "batch_per_gpu": B, "width": N_WIDTH},
                     cpu=lambda: cpu_commit_sample(ctx.bases_h, 16 * cores, cores) + (f"{16 * cores} of the {B} vectors",))
        else:
            def step():
                eng.commit_batch_dev(key, a, N_WIDTH, B, C)
                eng.kzg_open_batch_dev(key, a, N_WIDTH, z, B, pf, y)

            def step_e2e():   # commit + open of the same vectors: one upload of the rows
                check(L.vkzg_kzg_commit_open_batch(eng._ctx, kid, hp(a_h), ctypes.c_uint32(N_WIDTH), ctypes.c_uint32(0), hp(z_h),
                                                   ctypes.c_uint64(B), hp(C_h), hp(L_h), hp(y_h)), "commit+open")

            def checker():
                orc = _orc()
                Cn, pn, yn, an, zn = C.cpu().numpy(), pf.cpu().numpy(), y.cpu().numpy(), a_h.numpy(), z_h.numpy()
                for i in samples:
                    epf, ey, ok = orc.kzg_prove(ctx.bases_h[:N_WIDTH], an[i], zn[i])
                    _need(ok and (pn[i] == epf).all() and (yn[i] == ey).all(), f"opening {i} differs from the oracle")
                    _need((Cn[i] == orc.commit_batch(ctx.bases_h[:N_WIDTH], an[i:i + 1])[0]).all(), f"commit {i} differs from the oracle")
                _need((C_h.numpy() == Cn).all() and (L_h.numpy().reshape(-1)[: B * 64].reshape(B, 64) == pn).all() and (y_h.numpy() == yn).all(),
                      "e2e outputs differ from the device-resident outputs")
                return {"oracle_samples": len(samples), "e2e_equals_device": True}
            madds = 2 * N_WIDTH * W
            h2d, d2h = a_h.numel() + z_h.numel(), 2 * C_h.numel() + y_h.numel()
            w.update(metric="kzg_commit_and_open_per_s", unit="openings/s",
                     cfg={"workload": "configs[0] at batch: KZG commit + single-point open, width 256, batch 2^14 per GPU", "batch_per_gpu": B},
                     cpu=lambda: cpu_kzg_sample(ctx.bases_h, 8 * cores, cores) + (f"{8 * cores} of the {B} vectors",))
        w.update(step=step, step_e2e=step_e2e, check=checker, madds=madds, units=B, h2d=h2d, d2h=d2h)

    elif wl == "msm":
        n = 1 << args.log2n
        per = n // world                                                        # point-range sharding (configs[3])
        gen.manual_seed(0x5EED0004 + rank)
        bases = make_points_dev(torch, eng, per, gen)
        t_key = time.perf_counter()
        key = eng.load_key_dev(bases, per, kind=ctx._lib.KEY_MSM)
        eng.sync()
        w["extra"]["key_load_s"] = round(time.perf_counter() - t_key, 2)
        w["extra"]["msm_table_gb"] = round(key.table_bytes / 1e9, 2)
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

        def checker():
            """a 5000-point prefix against the oracle's bucket method; e2e == device; under torchrun every rank's combined
            result equals the group sum of the gathered partial sums recomputed by the oracle"""
            orc = _orc()
            k = min(5000, per)
            pre = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
            eng.msm_dev(key, s, k, pre)
            eng.sync()
            _need((pre.cpu().numpy()[0] == orc.msm(bases[:k].cpu().numpy(), s_h.numpy()[:k], mode="pippenger")).all(),
                  f"MSM over the first {k} points differs from the oracle")
            res = {"oracle_prefix_points": k}
            if world > 1:
                acc = np.zeros(64, dtype=np.uint8)
                for p in allp.cpu().numpy():
                    acc = orc.g1_add(acc, p)
                _need((out.cpu().numpy()[0] == acc).all() and (out_h.numpy()[0] == acc).all(), "combined MSM != sum of the gathered partial sums")
                res["combine_checked_ranks"] = world
            else:
                _need((out_h.numpy()[0] == part.cpu().numpy()[0]).all(), "e2e output differs from the device-resident output")
                res["e2e_equals_device"] = True
            return res

        def cleanup():
            key.free()
        nsmp = 128 * cores
        msm_windows = int(round(key.table_bytes / 64 / per))                   # digits per scalar of this key (16 at c = 16, 20 at c = 13)
        w["extra"]["msm_windows"] = msm_windows
        w.update(step=step, step_e2e=step_e2e, check=checker, cleanup=cleanup, madds=msm_windows, units=per, h2d=s_h.numel(), d2h=64,
                 kernel="k_msm_bucket", scaling="strong" if world > 1 else "weak",
                 metric="msm_points_per_s", unit="points/s",
                 cfg={"workload": f"configs[3]: one KZG-commit MSM of 2^{args.log2n} points, point-range sharded over {world} GPU(s)"
                                  + (", partial sums combined by all_gather + group sum" if world > 1 else ""), "points_per_gpu": per},
                 cpu=lambda: cpu_msm_sample(bases[:nsmp].cpu().numpy(), nsmp, cores) + (f"{nsmp} of the 2^{args.log2n} terms (naive per-term double-and-add scales linearly)",))

    elif wl == "multiproof":
        m = args.batch or (1 << 12)                                             # configs[2]: 2^12 openings aggregated
        key = ctx.key257(c)
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

        def checker():
            orc = _orc()
            got = dict(D=D_h.numpy().copy(), L=L_h.numpy().copy(), R=R_h.numpy().copy(), tip=tip_h.numpy().copy(), y=yo_h.numpy().copy())
            _need(eng.multiproof_verify_ipa(key, C_h.numpy(), z_h, y_h.numpy(), got), "the multiproof of the timed region fails the device verifier")
            exp = orc.multiproof_prove("ipa", ctx.bases_h, N_WIDTH, f_h.numpy(), C_h.numpy(), z_h, y_h.numpy())
            _need(all((got[k] == exp[k]).all() for k in got), "the multiproof differs from the oracle's")
            return {"verified": 1, "oracle_samples": 1}
        reps = min(cores, 4)
        w.update(step=step, step_e2e=step_e2e, check=checker, units=1,
                 madds=(2 * N_WIDTH + 8 * 2 * (N_WIDTH // 2 + 1)) * W,          # D, E commits + the final IPA opening
                 h2d=f_h.numel() + m * (64 + 8 + 32), d2h=64 + 2 * 8 * 64 + 64,
                 metric="ipa_multiproofs_per_s", unit="multiproofs/s",
                 cfg={"workload": f"configs[2]: IPA multiproof aggregating {m} openings at width 256 (one multiproof per step per GPU; replicas across GPUs)",
                      "openings_per_multiproof": m},
                 cpu=lambda: cpu_multiproof_sample(ctx.bases_h, m, cores, reps=reps) + (f"{reps} multiproofs of {m} openings side by side",))

    elif wl == "mpbatch":
        m = 1 << 12                                                             # configs[2] at a batch: K multiproofs of 2^12 openings in one call
        K = args.batch or 64
        key = ctx.key257(c)
        gen.manual_seed(0x5EED3100 + rank)
        tot = K * m
        f = rand_fr_dev(torch, tot * N_WIDTH, gen).reshape(tot, N_WIDTH, 32)
        zi = torch.randint(0, N_WIDTH, (tot,), device="cuda", generator=gen)
        Cq = torch.empty((tot, 64), dtype=torch.uint8, device="cuda")
        eng.commit_batch_dev(key, f, N_WIDTH, tot, Cq)
        yq = f[torch.arange(tot, device="cuda"), zi].contiguous()
        f_h = pinned(torch, (tot, N_WIDTH, 32))
        f_h.copy_(f)
        C_h, y_h = Cq.cpu().contiguous(), yq.cpu().contiguous()
        z_h = zi.cpu().numpy().astype(np.uint64)
        me = np.full(K, m, dtype=np.uint64)
        D_h, L_h, R_h = pinned(torch, (K, 64)), pinned(torch, (K, 8, 64)), pinned(torch, (K, 8, 64))
        tip_h, yo_h = pinned(torch, (K, 32)), pinned(torch, (K, 32))
        kid = ctypes.c_uint32(key.id)
        zp, mp_ = z_h.ctypes.data_as(ctypes.c_void_p), me.ctypes.data_as(ctypes.c_void_p)

        def step():
            check(L.vkzg_multiproof_prove_batch_dev(eng._ctx, kid, ctypes.c_int32(0), ctypes.c_void_p(f.data_ptr()), hp(C_h), zp, hp(y_h), mp_,
                                                    ctypes.c_uint64(K), hp(D_h), hp(L_h), hp(R_h), hp(tip_h), hp(yo_h)), "multiproof_batch_dev")

        def step_e2e():
            check(L.vkzg_multiproof_prove_batch(eng._ctx, kid, ctypes.c_int32(0), hp(f_h), hp(C_h), zp, hp(y_h), mp_, ctypes.c_uint64(K),
                                                hp(D_h), hp(L_h), hp(R_h), hp(tip_h), hp(yo_h)), "multiproof_batch")

        def checker():
            """every multiproof of the batch through the device verifier; the first and the last against the oracle"""
            orc = _orc()
            Cn, yn, fn = C_h.numpy(), y_h.numpy(), f_h.numpy()
            for i in range(K):
                pr = dict(D=D_h.numpy()[i], L=L_h.numpy()[i], R=R_h.numpy()[i], tip=tip_h.numpy()[i], y=yo_h.numpy()[i])
                sl = slice(i * m, (i + 1) * m)
                _need(eng.multiproof_verify_ipa(key, Cn[sl], z_h[sl], yn[sl], pr), f"multiproof {i} of the batch fails the device verifier")
                if i in (0, K - 1):
                    exp = orc.multiproof_prove("ipa", ctx.bases_h, N_WIDTH, fn[sl], Cn[sl], z_h[sl], yn[sl])
                    _need(all((pr[k] == exp[k]).all() for k in pr), f"multiproof {i} differs from the oracle's")
            return {"verified": K, "oracle_samples": min(K, 2)}
        reps = min(cores, 4)
        w.update(step=step, step_e2e=step_e2e, check=checker, units=K,
                 madds=(2 * N_WIDTH + 8 * 2 * (N_WIDTH // 2 + 1)) * W,
                 h2d=f_h.numel() + tot * (64 + 8 + 32), d2h=K * (64 + 2 * 8 * 64 + 64),
                 metric="ipa_multiproofs_per_s", unit="multiproofs/s",
                 cfg={"workload": f"configs[2] in bulk: {K} IPA multiproofs of {m} openings each (width 256) per call (vkzg_multiproof_prove_batch)",
                      "openings_per_multiproof": m, "multiproofs_per_call": K},
                 cpu=lambda: cpu_multiproof_sample(ctx.bases_h, m, cores, reps=reps) + (f"{reps} multiproofs of {m} openings side by side",))

    elif wl == "tree":
        nkeys = args.batch or (1 << 20)                                         # configs[4]: bulk insert of 2^20 keys
        from verkle_kzg_b200.tree import build_levels, NativeVerkleTree
        from verkle_kzg_b200.sharding import split_range
        key = ctx.key257(c)
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
                    n=cn, t=len(lv["slot"])) for lv, cn in zip(levels, counts)]
        nodes = torch.empty((total_nodes, 64), dtype=torch.uint8, device="cuda")
        allp = torch.empty((world, 64), dtype=torch.uint8, device="cuda")
        root = torch.empty((1, 64), dtype=torch.uint8, device="cuda")
        last = {}

        def step():
            off = 0
            for d in dev:
                eng.tree_level_dev(key, d["rp"], d["n"], d["sl"] if d["t"] else None, d["ch"] if d["t"] else None,
                                   d["li"] if d["t"] else None, d["t"], nodes, nodes[off:])
                off += d["n"]
            if world > 1:   # one exchange at the root: per-rank partial roots (disjoint child slots) are added
                dist.all_gather_into_tensor(allp, nodes[total_nodes - 1:total_nodes])
                eng.g1_sum_dev(allp, world, root)

        kk, vv = keys_all[sel], vals_all[sel]

        def step_e2e():
            # configs[4] end to end: bulk insert into the native host tree (libvkzg, Node::insert semantics) + recommit to the root
            t = NativeVerkleTree(32, 256)
            t.insert_many(kk, vv)
            last["root"] = t.commitment(eng, key)
            t.close()

        def checker():
            """the level-list path (value) and the native host tree (e2e) give the same root; a 300-key tree equals the oracle's"""
            orc = _orc()
            _need((nodes[total_nodes - 1].cpu().numpy() == last["root"]).all(), "level-list root != native-tree root")
            t = NativeVerkleTree(32, 256)
            t.insert_many(kk[:300], vv[:300])
            r300 = t.commitment(eng, key)
            t.close()
            _need((r300 == orc.tree_commit(ctx.bases_h[:N_WIDTH], kk[:300], vv[:300])).all(), "300-key tree root differs from the oracle")
            return {"roots_equal": True, "oracle_samples": 1}
        terms = sum(len(lv["slot"]) for lv in levels)
        # e2e traffic of vkzg_tree_commit: extensions travel compact (32-byte stem + unit + 32-byte value; the device writes
        # their CSR rows), internal levels as row_ptr + (slot, child) per term; every node commitment is cached back
        h2d = counts[1] * 65 + sum(4 * (cn + 1) + 6 * len(lv["slot"]) for lv, cn in zip(levels[2:], counts[2:]))
        w.update(step=step, step_e2e=step_e2e, check=checker, units=int(sel.sum()), scaling="strong" if world > 1 else "weak",
                 madds=terms * W / max(1, int(sel.sum())),                      # upper bound: zero digits are skipped at run time
                 h2d=h2d, d2h=64 * sum(counts[1:]), metric="verkle_tree_commit_keys_per_s", unit="keys/s",
                 cfg={"workload": f"configs[4]: verkle tree of {nkeys} random 32-byte keys, every node recommitted level by level up to the root",
                      "keys_this_rank": int(sel.sum()), "nodes_per_level": counts, "terms": terms, "host_flatten_seconds_not_timed": round(host_build_s, 1)},
                 cpu=lambda: cpu_tree_sample(ctx.bases_h, 256, cores) + (f"{cores} independent trees of 256 keys (node commitments by per-term double-and-add)",))
    else:
        raise SystemExit(f"unknown workload {wl}")
    if wl != "msm":
        w["extra"]["window_bits"] = c
        w["extra"]["window_table_gb"] = round(ctx._key.table_bytes / 1e9, 1)
    return w


def measure(ctx, wl, c, steps, warmup, sampler=None, cpu=True):
    """one workload -> its bench line (dict)"""
    torch, dist, world, rank, eng, args = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.eng, ctx.args
    w = build_workload(ctx, wl, c)
    step, step_e2e = w["step"], w["step_e2e"]
    cfg = w["cfg"]
    cfg.update(w["extra"])
    cfg["l2_policy"] = "tables (GBs) and inputs (>= 128 MB) exceed the 126 MB L2; no flush needed"
    # ---- device-resident timing (the `value`): W warm-up steps, K timed steps
    l0 = eng.launches
    ms, clocks = timed_steps(torch, dist, world, step, steps, warmup, sampler)
    launches = (eng.launches - l0) * steps // (steps + warmup)
    # ---- the dominant kernel's own launch durations (roofline): every launch bracketed by a CUDA event pair on its stream.
    #      Taken in a second pass of K steps with the IPA half-batches on ONE stream, so that a bracket times one kernel alone
    #      (in the timed region above the two half-batches interleave on two streams and brackets would overlap).
    eng.set_option(eng.OPT_IPA_TWO_STREAMS, 0)
    step()
    eng.kernel_timing(True)
    ms_k, _ = timed_steps(torch, dist, world, step, steps, 0)
    kn, kms = eng.kernel_timing_read()
    eng.kernel_timing(False)
    eng.set_option(eng.OPT_IPA_TWO_STREAMS, 1)
    units_per_step = w["units"]
    if world > 1:
        tu = torch.tensor([units_per_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tu)
        units_all = float(tu.item())
    else:
        units_all = float(units_per_step)
    value = units_all * steps / (ms * 1e-3)
    # ---- one more device-resident step so that the buffers the check reads hold a two-stream result (what was timed)
    step()
    torch.cuda.synchronize()
    # ---- end to end through the host-pointer C ABI
    # (2 warm-up steps: the host-pointer path has its own scratch and staging buffers to put into the stream-ordered pool; two
    #  timed repetitions of K/2 steps, the faster one reported: a one-off pool growth or a neighbour's PCIe burst inside a
    #  2-step window otherwise shows up as a 30 % outlier — seen once in ~20 runs; both figures are in the line)
    e2e_steps = max(1, steps // 2)
    reps = [timed_steps(torch, dist, world, step_e2e, e2e_steps, 2)[0], timed_steps(torch, dist, world, step_e2e, e2e_steps, 0)[0]]
    ms_e2e = min(reps)
    e2e_value = units_all * e2e_steps / (ms_e2e * 1e-3)
    # ---- check what the timed region produced
    checked = None
    if not args.no_check:
        try:
            checked = w["check"]()
            checked["ok"] = True
        except CheckFailed as e:
            checked = {"ok": False, "error": str(e)}
    if world > 1 and checked is not None:
        t = torch.tensor([0.0 if checked["ok"] else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        if t.item() > 0 and checked["ok"]:
            checked = {"ok": False, "error": "the check failed on another rank"}
        checked["ranks"] = world

    madds = w["madds"]
    macs_per_launch_set = units_per_step * madds * FQ_MUL_PER_MADD * MAC32_PER_FQ_MUL
    achieved = macs_per_launch_set * steps / (kms * 1e-3) / 1e12 if kms > 0 else None
    peak = ctx.peak
    # The same kernel against the HBM roofline (algorithmic bytes = one 64-byte table point per addition): the gather
    # stream is a few per cent of the measured copy bandwidth, i.e. the kernel is not memory-bound.
    hbm_peak, hbm_src = measured_hbm_peak()
    hbm_view = None
    if kn and kms:
        gbs = (units_per_step * madds * steps) * 64.0 / (kms * 1e-3) / 1e9
        hbm_view = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src}
    # SURVEY.md section 8d's per-unit figures (8-bit windows, Jacobian 11-mul additions): the work the REFERENCE algorithm
    # would need, not what this kernel executes (wider windows need fewer additions), so its "frac" can exceed 1.
    survey_mac = {"ipa": 12.3e6 + 98.9e6, "commit": 12.3e6, "kzg": 12.3e6 + 0.17e6,
                  "msm": {16: 35.4e3, 18: 31.0e3, 20: 26.1e3}.get(getattr(args, "log2n", 20))}.get(wl)
    survey_acct = None
    if survey_mac and kms:
        a = units_per_step * survey_mac * steps / (kms * 1e-3) / 1e12
        survey_acct = {"mac32_per_unit": survey_mac, "achieved": a, "frac": a / peak if peak else None,
                       "note": "reference-algorithm work per unit; > 1 means fewer operations were executed than that algorithm needs"}
    line = {
        "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
        "dtype": "u32 limbs (254-bit modular integers)", "data": "synthetic", "config": cfg,
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": w["unit"], "h2d_bytes_per_step": int(w["h2d"]), "d2h_bytes_per_step": int(w["d2h"]),
                "timing": f"faster of 2 repetitions of {e2e_steps} steps after 2 warm-up steps, CUDA events, max over ranks",
                "repetitions": [units_all * e2e_steps / (t * 1e-3) for t in reps]},
        "checked": checked,
        "roofline": {
            "bound": "int32",
            "bound_note": "integer multiply pipe (IMAD): 254-bit modular arithmetic is neither hbm- nor tensor-bound; the hbm view of the same kernel is under hbm_view",
            "kernel": w["kernel"],
            "achieved": achieved, "peak": peak, "unit": "TMAC32/s", "frac": (achieved / peak) if achieved else None,
            "peak_source": "measured live: dependency-free mad.wide.u32 chains on all SMs (vkzg_probe_imad_dev); MEASURED_PEAKS.json has no integer figure",
            "work_model": f"{madds} mixed additions per unit x {FQ_MUL_PER_MADD} Fq-mul x {MAC32_PER_FQ_MUL} MAC32",
            # what the kernel actually issues per mixed addition: 6 general products (136), 2 dedicated squares (108) and one
            # fused pair sharing its reduction (200) = 1232 multiply-accumulates for the 1360 of the accounting above
            "executed_mac32_per_addition": 6 * 136 + 2 * 108 + 200,
            "survey_accounting": survey_acct,
            "kernel_launches_timed": kn, "kernel_ms_total": kms, "kernel_share_of_step": kms / ms_k if ms_k else None,
            "one_stream_ms_per_step": ms_k / steps,
            # DRAM bytes per launch: NOT measured in this run — 128 B per table addition as ncu measured it on this kernel
            # (dram__bytes_read.sum + dram__bytes_write.sum of one --set full capture under profiles/: 1.785 GB for a launch of
            # 13.7 M additions; every 64-byte point is fetched at 128-byte granularity) x additions per launch
            "traffic": (units_per_step * madds * steps / kn) * 128.0 if kn else None,
            "traffic_source": "modelled: 128 B per table addition from the ncu --set full capture under profiles/, scaled to this launch size (not measured live)",
            "algorithmic_bytes_per_launch": (units_per_step * madds * steps / kn) * 64.0 if kn else None,
            "hbm_view": hbm_view,
            "ncu": "profiles/: sm__pipe_fmaheavy_cycles_active and DRAM bytes of k_fixed_base_msm / k_msm_bucket",
        },
    }
    if cpu and not args.no_cpu_baseline and world == 1 and rank == 0:  # the CPU baseline is reported at N = 1 only
        dt, u, sample = w["cpu"]()
        line["cpu_baseline"] = {"value": u / dt, "unit": w["unit"], "cores": cpu_cores(), "kind": "port", "seconds": round(dt, 2),
                                "sample": sample + "; restated reference algorithm (C++ oracle), not the arkworks binary"}
    w["cleanup"]()
    del w
    torch.cuda.empty_cache()
    return line


def single_call_latencies(ctx, c):
    """configs[0] as the reference's own criterion bench runs it (benches/kzg.rs:61-75, benches/ipa.rs:64-109): ONE vector per
    call from host buffers, wall clock of the C-ABI call — the latency a drop-in caller of `commit` + `prove` sees — next to
    the oracle's single-thread time for the same vector (the reference is single-threaded there).  Reported, not a metric."""
    torch, eng, L = ctx.torch, ctx.eng, ctx.L
    check = ctx._lib.check
    orc = _orc()
    key = ctx.key257(c)
    kid = ctypes.c_uint32(key.id)
    rng = np.random.default_rng(0xB1)
    a_h, z_h = pinned(torch, (1, N_WIDTH, 32)), pinned(torch, (1, 32))
    a_h.copy_(torch.from_numpy(orc.rand_fr_buf(rng, N_WIDTH).reshape(1, N_WIDTH, 32)))
    z_h.copy_(torch.from_numpy(orc.fr_to_buf([77])))
    C_h, P_h, R_h = pinned(torch, (1, 64)), pinned(torch, (1, 8, 64)), pinned(torch, (1, 8, 64))
    tip_h, y_h = pinned(torch, (1, 32)), pinned(torch, (1, 32))
    one = ctypes.c_uint64(1)

    def kzg():
        check(L.vkzg_kzg_commit_open_batch(eng._ctx, kid, hp(a_h), ctypes.c_uint32(N_WIDTH), ctypes.c_uint32(0), hp(z_h), one, hp(C_h), hp(P_h),
                                           hp(y_h)), "commit+open")

    def ipa():
        check(L.vkzg_ipa_commit_prove_batch(eng._ctx, kid, hp(a_h), hp(z_h), one, hp(C_h), hp(P_h), hp(R_h), hp(tip_h), hp(y_h)), "commit+prove")

    out = {}
    an, zn = a_h.numpy(), z_h.numpy()
    for name, fn in (("kzg_commit_open", kzg), ("ipa_commit_prove", ipa)):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        eC = orc.commit_batch(ctx.bases_h[:N_WIDTH], an)[0]
        t0 = time.perf_counter()
        if name == "kzg_commit_open":
            orc.commit_batch(ctx.bases_h[:N_WIDTH], an, nthreads=1)
            epf, ey, ok = orc.kzg_prove(ctx.bases_h[:N_WIDTH], an[0], zn[0])
            good = ok and (P_h.numpy().reshape(-1)[:64] == epf).all() and (y_h.numpy()[0] == ey).all() and (C_h.numpy()[0] == eC).all()
        else:
            orc.commit_batch(ctx.bases_h[:N_WIDTH], an, nthreads=1)
            eL, eR, etip, ey = orc.ipa_prove(ctx.bases_h, N_WIDTH, an[0], eC, zn[0])
            good = ((P_h.numpy()[0] == eL).all() and (R_h.numpy()[0] == eR).all() and (tip_h.numpy()[0] == etip).all()
                    and (y_h.numpy()[0] == ey).all() and (C_h.numpy()[0] == eC).all())
        cpu_ms = (time.perf_counter() - t0) * 1e3
        out[name] = {"ms_median": round(float(np.median(ts)) * 1e3, 4), "ms_min": round(min(ts) * 1e3, 4), "cpu_ms_1thread": round(cpu_ms, 2),
                     "matches_oracle": bool(good)}
    out["note"] = ("one width-256 vector per call, host buffers, wall clock around the C-ABI call (B = 1: latency-bound, DESIGN.md "
                   "section 4); cpu = the oracle's restatement on ONE host thread, as the reference runs it")
    return out


def condensed(line):
    """what an `also` entry keeps of a full line"""
    r = line["roofline"]
    out = {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "steps": line["steps"],
           "warmup": line["warmup"], "scaling": line["scaling"], "gpu_launches": line["gpu_launches"],
           "e2e": {k: line["e2e"][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")},
           "roofline": {"kernel": r["kernel"], "frac": r["frac"], "achieved": r["achieved"], "peak": r["peak"], "unit": r["unit"],
                        "kernel_share_of_step": r["kernel_share_of_step"], "work_model": r["work_model"]},
           "checked": line["checked"], "config": line["config"]}
    if "cpu_baseline" in line:
        out["cpu_baseline"] = line["cpu_baseline"]
    return out


def run_native(args):
    ctx = Ctx(args)
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    c = ctx.window_bits
    line = measure(ctx, args.workload, c, args.steps, args.warmup, sampler)
    failed = line["checked"] is not None and not line["checked"]["ok"]
    if args.workload == "ipa" and not args.no_also and not args.batch:
        also = {}
        k, wu = min(args.steps, 10), 3
        plan = [("commit_w256", "commit", c), ("kzg_open", "kzg", c), ("multiproof_2p12", "multiproof", c), ("multiproof_2p12_x64", "mpbatch", c),
                ("tree_2p20", "tree", c),
                (f"msm_2p{args.log2n}", "msm", c)]
        if c != 16:   # the library's default window width beside the bench's pick
            plan += [("ipa_c16", "ipa", 16), ("commit_w256_c16", "commit", 16)]
        for name, wl, cc in plan:
            try:
                sub = measure(ctx, wl, cc, k, wu, None, cpu=not name.endswith("_c16"))
                also[name] = condensed(sub)
                failed = failed or (sub["checked"] is not None and not sub["checked"]["ok"])
            except CheckFailed as e:   # (measure() catches these itself; kept for safety)
                also[name] = {"error": str(e)}
                failed = True
        line["also"] = also
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["single_call"] = single_call_latencies(ctx, c)
                failed = failed or not all(v["matches_oracle"] for v in line["single_call"].values() if isinstance(v, dict))
            except Exception as e:   # a reported extra must not take the bench line down
                line["single_call"] = {"error": repr(e)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if failed:
        sys.stderr.write("bench.py: OUTPUT CHECK FAILED — see `checked` in the JSON line\n")
        sys.exit(3)


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
