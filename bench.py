#!/usr/bin/env python3
"""bench.py -- polymul/s of the batched NTT polynomial multiplier (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

The default run (workload c2) also measures, inside the same JSON line, `sustained` (c2 back to
back for >= 1 s with the clocks sampled throughout), `workloads` {c3, c4, c5} (every other GPU
configuration of BASELINE.json: value, ms per step, both roofline fractions, clocks, parity) and
`strong_scaling` {c3, c4} (ONE fixed batch of 2^20 / 2^18 rows sliced over the N ranks by
nttb200_shard_bounds, as BASELINE.json configs[2]/[3] state them).

One "step" = one pass of the hot path (fused NTT -> pointwise -> INTT kernel) over one batch
of synthetic polynomial pairs.  Default workload = BASELINE.json configs[1]: batch 2^16 at
the reference default (n=256, q=12289).  With N>1 (torchrun, one process per GPU) every rank
runs the same per-GPU batch on its own rows -- no data-path collective (the products share
nothing), torch.distributed only provides the barrier and the max-over-ranks of the time.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the same
metric through the host-buffer C-ABI call (pinned host memory, H2D and D2H inside the timed
region); `roofline` is the fused kernel against the measured HBM peak; `int_roofline` is the
same kernel against the integer-multiply issue rate measured live; `cpu_baseline` is the
reference's own C code on this box's host cores (N=1 only).

`--impl reference` times the reference's CPU implementation instead (oracle/_ref compiled
from the unmodified sources; the oracle port for (n,q) the reference cannot run).
"""
from __future__ import annotations

import argparse
import importlib
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "ntt-based-polynomial-multiplier-fpga_b200"

WORKLOADS = {
    # name: (n, q, psi, log2 batch, description)
    "c2": (256, 12289, 1002, 16, "BASELINE configs[1]: batch 2^16 polymuls at the reference default (n=256,q=12289)"),
    "c3": (256, 7681, 0, 20, "BASELINE configs[2] (Kyber-like): n=256 q=7681 (3329 has no 512-th root), batch 2^20"),
    "c3c": (256, 3329, 0, 20, "BASELINE configs[2], literal q=3329: cyclic product mod x^256-1 (the psi-free surface), batch 2^20"),
    "c4": (1024, 12289, 0, 18, "BASELINE configs[3] (Falcon/NewHope-like): n=1024 q=12289 batch 2^18"),
    "c5h": (65536, 998244353, 0, 10, "variant of configs[4] with a 30-bit prime (119*2^23+1): HARVEY class, 6-instruction butterflies"),
    "c5": (65536, 2013265921, 0, 10, "BASELINE configs[4]: n=2^16 q=2013265921 batch 2^10, multi-pass"),
}
# DRAM bytes (read + write) per launch of the dominant kernel, from `ncu --set full` captures
# committed under profiles/ (a profiler run is never a bench value; this is the traffic only)
TRAFFIC_NCU = {"c2": 134266624 + 37117952, "c4": 2147526000 + 1041490000}
TRAFFIC_SRC = {"c2": "profiles/r2_c2_splant_final_ncu_full.txt: dram__bytes_read.sum 134.26 MB + "
                     "dram__bytes_write.sum 37.12 MB per launch (algorithmic 201.3 MB; part of c stays dirty in L2)",
               "c4": "profiles/r2_c4_splant_wide_final_ncu_full.txt: 2.148 GB read + 1.041 GB written per launch (algorithmic 3.221 GB)"}
SEED = 0x4E545442323030
L2_BYTES = 126 * 1000 * 1000


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
# clocks: NVML sampling thread covering the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._timed = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((self._timed.is_set(), mhz))
                if self._timed.is_set():
                    for bit, name in self.REASONS.items():
                        if r & bit and name != "gpu_idle":
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def timed(self, on: bool):
        (self._timed.set if on else self._timed.clear)()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=2)
        timed = [m for t, m in self.samples if t] or [m for _, m in self.samples]
        return {"sm_mhz": statistics.median(timed) if timed else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples_in_timed_region": sum(1 for t, _ in self.samples if t)}


# --------------------------------------------------------------------------------------
# CPU side: the reference's own C code on the host cores
# --------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(n, q, psi, use_ref, variant):
    from oracle import loader
    _W["variant"] = variant
    _W["n"], _W["q"], _W["psi"] = n, q, psi
    if use_ref:
        _W["ref"] = loader.Reference()
    else:
        _W["orc"] = loader.Oracle()


def _cpu_worker_loop(args):
    """Timed loop (cpu_baseline leg): seconds of work on a private slice, returns (rate, calls)."""
    seed, rows, seconds = args[:3]
    variant = args[3] if len(args) > 3 and args[3] is not None else _W["variant"]
    flavour = args[4] if len(args) > 4 else None
    from oracle import loader
    O = loader.Oracle()
    n, q = _W["n"], _W["q"]
    a, b = O.random((rows, n), q, seed), O.random((rows, n), q, seed + 1)
    if "ref" in _W:
        ref = _W["ref"]
        if flavour:
            ref = _W.setdefault("ref_" + flavour, loader.Reference(flavour))
        return ref.bench_loop(a, b, variant, seconds)
    return _W["orc"].bench_loop(n, q, a, b, variant, seconds, _W["psi"])


def _cpu_worker_step(args):
    """One slice of one step (--impl reference leg): the slice is multiplied `reps` times."""
    seed, rows, reps = args
    key = ("data", seed, rows)
    from oracle import loader
    if key not in _W:
        O = loader.Oracle()
        _W[key] = (O.random((rows, _W["n"]), _W["q"], seed), O.random((rows, _W["n"]), _W["q"], seed + 1))
    a, b = _W[key]
    t0 = time.perf_counter()
    for _ in range(reps):
        if "ref" in _W:
            c = _W["ref"].product(a, b, _W["variant"])
        else:
            c = _W["orc"].product(_W["n"], _W["q"], a, b, _W["variant"], _W["psi"])
    return time.perf_counter() - t0, int(c[0, 0])


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_pool(n, q, psi, cores):
    from oracle import loader
    use_ref = (n, q) == (256, 12289) and loader.reference_available()
    variant = loader.REF_RED_CT if use_ref else (loader.PRODUCT_MERGED if (q - 1) % (2 * n) == 0 else loader.PRODUCT_CYCLIC)
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(n, q, psi, use_ref, variant))
    kind = "reference" if use_ref else "port"
    what = ("ntt_red256_product1 (optimized CT, Longa-Naehrig) from oracle/_ref, gcc -O3, unmodified sources"
            if use_ref else "oracle port of the merged CT-fwd/GS-inv pipeline (ntt_oracle.c)")
    return pool, kind, what


def cpu_baseline(n, q, psi, seconds=4.0):
    cores = host_cores()
    pool, kind, what = cpu_pool(n, q, psi, cores)
    rows = max(1, min(4096, (1 << 20) // n))
    variants = {}
    try:
        res = pool.map(_cpu_worker_loop, [(SEED + 17 * i, rows, seconds) for i in range(cores)])
        if kind == "reference":   # secondary rows of BASELINE.md section 3, 1 s each, all cores
            from oracle import loader
            for name, var, flav in (("ntt_red256_product4 (optimized GS)", loader.REF_RED_GS, None),
                                    ("ntt256_product1 (CT)", loader.REF_CT, None),
                                    ("ntt256_product4 (GS)", loader.REF_GS, None),
                                    ("merged CT-fwd/GS-inv pipeline", loader.REF_MERGED, None),
                                    ("ntt_red256_product1, as-documented flag-less build (-O0)", loader.REF_RED_CT, "O0")):
                try:
                    r2 = pool.map(_cpu_worker_loop, [(SEED + 17 * i, rows, 1.0, var, flav) for i in range(cores)])
                    variants[name] = sum(r for r, _ in r2)
                except Exception as ex:
                    variants[name] = repr(ex)
    finally:
        pool.close()
        pool.join()
    total = sum(r for r, _ in res)
    return {"value": total, "unit": "polymul/s", "cores": cores, "kind": kind,
            "per_core": total / cores, "other_variants_polymul_per_s": variants,
            "sample": f"{what}; one process per core, each looping over a private slice of {rows} random "
                      f"polymuls for {seconds:.0f} s (operand restore excluded, as time_testing256.c:175-185)"}


def run_reference_arm(args, n, q, psi, logb, desc, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    pool, kind, what = cpu_pool(n, q, psi, cores)
    full = 1 << logb
    # bounded sample per step: ~0.3 s of all-core work at ~3e5 polymul/s/core (n=256), as `reps`
    # passes over a per-core slice of at most 4096 rows (the process-pool hand-off is then noise)
    est_rate = cores * 3.0e5 * (256 * 8) / (n * max(1, n.bit_length() - 1))
    per = max(1, min(4096, (1 << 20) // n, full // cores if full >= cores else 1))
    step_s = min(0.3, max(0.03, 90.0 / max(1, args.steps + args.warmup)))   # whole run <= ~1.5 min
    reps = max(1, int(est_rate * step_s / (per * cores)))
    sample = per * cores * reps
    jobs = [(SEED + 31 * i, per, reps) for i in range(cores)]
    try:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_worker_step, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker_step, jobs)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "polymul/s", "value": value, "unit": "polymul/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc, "n": n, "q": q, "batch_per_gpu": full,
                   "l2": "inputs larger than L2 (CPU arm: each process loops over its own slice)"},
        "setup": {"batch_per_step": sample, "note": "CPU arm: host cores only, no GPU, no host<->device copies; each "
                  "step is a bounded sample of the workload (cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": "polymul/s", "cores": cores, "kind": kind,
                         "sample": f"{what}; each step = {sample} polymuls ({reps} passes over {per} rows "
                                   f"per process, {cores} processes = one per core; the workload has {full}), "
                                   f"wall clock including the operand restore the reference needs"},
        "e2e": {"value": value, "unit": "polymul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def bind_near_gpu(local: int, world: int):
    """N > 1, host-buffer calls: every rank narrows, stages and widens its rows with a pool of host
    threads into pinned memory, and eight ranks share one box.  Keep each rank's threads -- and, by
    first touch, its pinned pages -- on the CPUs NVML reports as local to its GPU, and split those
    CPUs between the ranks that share them.  Returns (cpus of this rank, threads for its pool) or None
    when NVML has nothing to say (then the ranks simply split all cores evenly)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        allowed = os.sched_getaffinity(0)
        words = (max(allowed) // 64) + 1

        def near(i):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            return frozenset(64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1) & allowed
        mine = near(local)
        sharers = [i for i in range(world) if near(i) == mine]
        if len(mine) < 2 * len(sharers):
            return None                                   # too few local cores to be worth a binding
        cpus = sorted(mine)
        k = sharers.index(local)
        part = cpus[len(cpus) * k // len(sharers): len(cpus) * (k + 1) // len(sharers)]
        os.sched_setaffinity(0, part)
        return part, len(part)
    except Exception:
        return None


def load_fixture():
    """The reference's coefficient files (row 8 of the (256, 12289) batch); the committed
    copy under tests/golden (a fixture file, not the oracle)."""
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_256_12289.npz"))
        return g["fixture_a"], g["fixture_b"]
    except Exception:
        return None


def slots_per_polymul(n: int, plantard: bool, signed: bool = False) -> int:
    """fmaheavy issue slots per product as the kernels are written: Shoup butterfly = IMAD.HI (2 slots:
    measured half rate) + 2 IMAD; pointwise Montgomery = 2 IMAD.HI + 2 IMAD; the n^-1 scaling costs one
    extra multiplication on the sum branch of the last stage.  Plantard (q <= 12385): butterfly 3,
    pointwise 4, scale 3; at n <= 256 butterfly and scale 2 (IMAD, SHF, IMAD, SHF).  Signed Plantard
    kernels (ntt_small_splant.cuh, ntt_splant_wide.cuh: the default): L - 3 stages per transform with 2
    multiplications per butterfly, 4 in the butterflies of the last inverse stage, and per group of eight
    coefficients 8 Barrett steps (2 each), 64 + 7 raw products and 15 reductions (2 each): n (3 L + 6.625)."""
    bflies = 3 * (n // 2) * (n.bit_length() - 1)
    L = n.bit_length() - 1
    if plantard and signed:
        return 2 * n * (L - 3) + (n // 8) * 117 + n * (L - 4) + 2 * n
    if plantard and n <= 256:
        return 2 * bflies + 4 * n + 2 * (n // 2)
    if plantard:
        return 3 * bflies + 4 * n + 3 * (n // 2)
    return 4 * bflies + 6 * n + 4 * (n // 2)


def check_rows(mod, n, q, cyclic, a, b, c, rows, fixture_c=None):
    """Rows of the timed output against the CPU checker: the compiled unmodified reference
    (oracle/_ref: optimized CT and plain GS variants) at (256, 12289), the oracle port elsewhere."""
    from oracle import loader
    import torch
    ti = torch.as_tensor(rows, device=a.device)
    ha, hb, hc = a[ti].cpu().numpy(), b[ti].cpu().numpy(), c[ti].cpu().numpy()
    if (n, q) == (256, 12289) and loader.reference_available():
        R = loader.Reference()
        ok = all(bool((R.product(ha, hb, v) == hc).all()) for v in (loader.REF_RED_CT, loader.REF_GS))
        return ok, "oracle/_ref (ntt_red256_product1, ntt256_product4)"
    from concurrent.futures import ThreadPoolExecutor
    O = loader.Oracle()
    variant = loader.PRODUCT_CYCLIC if cyclic else loader.PRODUCT_MERGED
    O.plan(n, q, 0)
    th = max(1, min(host_cores(), len(rows) // 2))
    cuts = [len(rows) * i // th for i in range(th + 1)]
    with ThreadPoolExecutor(th) as ex:
        parts = list(ex.map(lambda i: O.product(n, q, ha[cuts[i]:cuts[i + 1]], hb[cuts[i]:cuts[i + 1]], variant), range(th)))
    return bool((np.concatenate(parts) == hc).all()), "oracle port (ntt_oracle.c, merged CT-fwd/GS-inv)"


class Workload:
    """One BASELINE configuration resident on this rank's GPU: plan, rotating operand sets built
    from the SURVEY 8d generator, and the timing loops over it."""

    def __init__(self, mod, sh, torch, name, dev, rank, world, rows_per_gpu=None, row_offset=None):
        self.mod, self.sh, self.torch, self.name, self.dev, self.rank, self.world = mod, sh, torch, name, dev, rank, world
        self.n, self.q, self.psi, logb, self.desc = WORKLOADS[name]
        self.full_batch = 1 << logb
        self.batch = rows_per_gpu if rows_per_gpu is not None else self.full_batch
        self.cyclic = (self.q - 1) % (2 * self.n) != 0          # q=3329, n=256: no 512-th root of unity
        self.plan = mod.Plan(self.n, self.q, self.psi, cyclic=self.cyclic)
        self.plantard = "plantard" in self.plan.describe()
        self.signed = "plantard-signed" in self.plan.describe()
        row_bytes = self.n * 4
        # distinct buffer sets so that successive steps never find their inputs in the 126 MB L2; a
        # stream of fresh batches is also what the call is for, so rotate over 8 sets (24 GiB at most):
        # consecutive steps then share no buffer, as in a pipeline that never reuses one in flight
        self.sets = max(2, -(-3 * L2_BYTES // (3 * self.batch * row_bytes)) + 1)
        want = int(os.environ.get("NTTB200_BENCH_SETS", "8"))
        self.sets = max(self.sets, min(want, (24 << 30) // (3 * self.batch * row_bytes)))
        cfg = {"c2": 2, "c3": 3, "c3c": 3, "c4": 4, "c5": 5, "c5h": 5}[name]
        off = row_offset if row_offset is not None else rank * self.batch
        fixture = load_fixture()
        self.bufs = []
        for k in range(self.sets):
            # set 0 is the SURVEY 8d batch itself (splitmix64 stream of config `cfg`, edge rows 0..7,
            # the reference's coefficient files in row 8); the other sets continue the same stream
            a, b = mod.inputs.survey_batch(self.n, self.q, self.batch, cfg, device=dev, fixture=fixture,
                                           row_offset=off + k * (self.full_batch * max(1, world)))
            self.bufs.append((a, b, torch.empty_like(a)))
        self.stream = torch.cuda.current_stream().cuda_stream
        self.ptrs = [(c.data_ptr(), a.data_ptr(), b.data_ptr()) for a, b, c in self.bufs]
        self.alg_bytes = 12 * self.n * self.batch               # read a, read b, write c (int32 API)
        self.launches_per_step = None
        # setup, not a step: every buffer set goes through the call once, so that no timed step is the
        # first ever to touch its result buffer (with W = 5 warm-up steps over 8 sets three of them would
        # be: measured 1 296 vs 1 327 M polymul/s for the K = 20 region, tools/k20_probe.py)
        for i in range(self.sets):
            self.step(i)
        torch.cuda.synchronize()

    def step(self, i):
        c, a, b = self.ptrs[i % self.sets]
        self.plan.polymul_dev(c, a, b, self.batch, self.stream)

    def timed(self, steps, warmup, local, clocks=True):
        """K steps between two events, barrier + synchronize on both sides, max over ranks."""
        torch, sh = self.torch, self.sh
        sampler = ClockSampler(local) if clocks else None
        if sampler:
            sampler.start()
        for i in range(warmup):
            self.step(i)
        self.launches_per_step = self.mod.last_launch_count()
        torch.cuda.synchronize()
        sh.barrier(local)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.timed(True)
        e0.record()
        for i in range(steps):
            self.step(i)
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            sampler.timed(False)
        sh.barrier(local)
        ms_local = e0.elapsed_time(e1)
        ms = sh.max_over_ranks(ms_local, self.dev)
        return ms, ms_local, (sampler.stop() if sampler else None)

    def per_launch_ms(self, steps):
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(steps, 50))]
        for i, (s, e) in enumerate(evs):
            s.record()
            self.step(i)
            e.record()
        torch.cuda.synchronize()
        return statistics.mean(s.elapsed_time(e) for s, e in evs)

    def parity(self, last_step, nrows):
        """The output of the last timed step against the CPU checker: edge rows 0..8, the last rows and
        `nrows` rows drawn over the batch."""
        a, b, c = self.bufs[last_step % self.sets]
        rng = np.random.default_rng(7)
        rows = np.unique(np.r_[0:min(9, self.batch), max(0, self.batch - 4):self.batch,
                               rng.integers(0, self.batch, nrows)])
        ok, checker = check_rows(self.mod, self.n, self.q, self.cyclic, a, b, c, rows)
        return ok, checker, int(rows.size)

    def fractions(self, rate_per_gpu, peak_gbs, imad_peak):
        n = self.n
        modmuls = 3 * (n // 2) * (n.bit_length() - 1) + n      # M
        slots = slots_per_polymul(n, self.plantard, self.signed)
        return {"hbm_frac": rate_per_gpu * 12 * n / 1e9 / peak_gbs,
                "imad_frac_survey_3_per_modmul": rate_per_gpu * 3 * modmuls / imad_peak if imad_peak else None,
                "imad_slot_frac_as_written": rate_per_gpu * slots / imad_peak if imad_peak else None}

    def close(self):
        self.plan.close()
        self.bufs = []
        self.torch.cuda.empty_cache()


def side_workload(mod, sh, torch, name, dev, rank, world, local, peak_gbs, imad_peak, budget_s=0.35,
                  rows_per_gpu=None, row_offset=None, parity_rows=None):
    """A non-headline BASELINE configuration, measured in the same run: value, time per step, both
    roofline fractions, clocks during its timed region, parity of its last output."""
    w = Workload(mod, sh, torch, name, dev, rank, world, rows_per_gpu=rows_per_gpu, row_offset=row_offset)
    try:
        ms1, _, _ = w.timed(2, 3, local, clocks=False)              # sizing run
        steps = int(max(5, min(400, budget_s * 1e3 / max(ms1 / 2, 1e-3))))
        ms, ms_local, clocks = w.timed(steps, 3, local)
        total_rows = sh.sum_over_ranks(float(w.batch), dev)
        value = total_rows * steps / (ms * 1e-3)
        if parity_rows is None:
            parity_rows = 4096 if w.n <= 1024 else 24
        ok, checker, nchecked = w.parity(steps - 1, parity_rows)
        rec = {"workload": w.desc, "n": w.n, "q": w.q, "rows_per_gpu": w.batch, "rows_total": int(total_rows),
               "value": value, "unit": "polymul/s", "steps": steps, "warmup": 3, "ms_per_step": ms / steps,
               "launches_per_step": w.launches_per_step, "plan": w.plan.describe(), "clocks": clocks,
               "parity_ok": ok, "parity_rows_checked": nchecked, "parity_checker": checker,
               "l2": f"{w.sets} rotating buffer sets ({w.sets * 3 * w.batch * w.n * 4 >> 20} MiB)"}
        rec.update(w.fractions(w.batch * steps / (ms_local * 1e-3), peak_gbs, imad_peak))
        if "arith=canon" in rec["plan"]:
            rec["butterfly_roof"] = butterfly_roof_canon(mod, w.n, w.batch * steps / (ms_local * 1e-3))
        return rec
    finally:
        w.close()


def butterfly_roof_canon(mod, n, rate_per_gpu):
    """31-bit moduli (CANON class): what the integer pipes allow BEFORE a byte moves.  The three butterflies of
    the large-n kernels are timed in isolation on every SM (nttb200_measure_int_peak 24 / 25 / 26: independent
    chains, no memory instruction), and a product is counted as they occur in it: per transform log2(n) stages
    of n/2 butterflies; forward, half the butterflies of three stages in four leave their results in [0, 2q),
    and so does the whole last stage of operand a (modarith.cuh: LAZYOUT) -- 13 of the 32 forward stages'
    worth; the 16 inverse stages are Gentleman-Sande.  The pointwise products (two IMAD.HI and two IMAD each)
    are counted at the measured IMAD.HI rate."""
    L = n.bit_length() - 1
    ct, ct_lazy, gs = (mod.measure_int_peak(k) for k in (24, 25, 26))
    hi = mod.measure_int_peak(1)
    lazy_stages = 2 * (L - L // 4) * 0.5 + 1          # three stages in four of both operands, half of each; + a's last
    canon_stages = 2 * L - lazy_stages
    sec = (n // 2) * (lazy_stages / ct_lazy + canon_stages / ct + L / gs) + n * 3 / hi
    return {"ct_canonical_per_s": ct, "ct_lazy_out_per_s": ct_lazy, "gs_per_s": gs,
            "polymul_per_s": 1.0 / sec, "frac": rate_per_gpu * sec,
            "note": "isolated butterfly rates x the product's butterfly counts: the roof of this arithmetic on the "
                    "integer pipes, memory instructions and launches not counted"}


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner and, at NCCL_DEBUG=INFO, its whole log to the C-level stdout): keep a private copy of the
    real stdout for the JSON line and point file descriptor 1 at stderr for everybody else, so
    nothing is lost and NCCL_DEBUG stays as the operator set it."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main() -> int:
    out = protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-workloads", action="store_true",
                    help="skip the c3/c4/c5 records and the strong-scaling record of the default run")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    args = ap.parse_args()
    n, q, psi, logb, desc = WORKLOADS[args.workload]
    batch = 1 << logb

    if args.impl == "reference":
        return run_reference_arm(args, n, q, psi, logb, desc, out)

    import torch
    mod = importlib.import_module(PKG)
    sh = importlib.import_module(PKG + ".sharding")
    rank, world, local = sh.env_rank_world()
    if not torch.cuda.is_available() or mod.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    mod.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sh.init_distributed("nccl")
    # the ranks of one box share its host cores: split them between the per-process pools that
    # narrow / widen the rows of the host-buffer call (csrc/hostwire.c), each rank next to its GPU
    binding = bind_near_gpu(local, world) if world > 1 and os.environ.get("NTTB200_BENCH_BIND", "1") != "0" else None
    os.environ.setdefault("NTTB200_HOST_THREADS",
                          str(binding[1] if binding else max(1, host_cores() // max(1, world))))
    warmup = max(3, args.warmup)
    peak_gbs, peak_src = measured_peaks()
    imad_peak = mod.measure_int_peak(0)
    imadhi_peak = mod.measure_int_peak(1)
    bfly_peak = mod.measure_int_peak(3)

    W = Workload(mod, sh, torch, args.workload, dev, rank, world)
    plan, bufs, sets, stream = W.plan, W.bufs, W.sets, W.stream
    row_bytes = n * 4
    ms, ms_local, clocks = W.timed(args.steps, warmup, local)
    rank_ms = sh.gather_over_ranks(ms_local, dev) if world > 1 else [ms_local]
    launches_per_step = W.launches_per_step
    value = world * batch * args.steps / (ms * 1e-3)
    parity_ok, parity_checker, parity_rows = W.parity(args.steps - 1, 4096 if n <= 1024 else 24)

    # per-launch duration of the dominant kernel, CUDA events on the launching stream
    launch_ms = W.per_launch_ms(args.steps)
    # a back-to-back stream hides launch gaps: the K-step region gives the better per-launch figure
    per_launch_ms = min(launch_ms, ms_local / args.steps) / max(1, launches_per_step)

    # sustained: the same step back to back for >= 1 s (the K-step region above can be as short as a
    # millisecond), with the clocks sampled every 2 ms throughout
    sust_steps = int(max(args.steps, min(200000, 1.25e3 / max(ms_local / args.steps, 1e-3))))
    ms_s, ms_s_local, clocks_s = W.timed(sust_steps, 3, local)
    sustained = {"value": world * batch * sust_steps / (ms_s * 1e-3), "unit": "polymul/s", "steps": sust_steps,
                 "seconds": ms_s * 1e-3, "ms_per_step": ms_s / sust_steps, "clocks": clocks_s,
                 "hbm_frac": batch * sust_steps / (ms_s_local * 1e-3) * 12 * n / 1e9 / peak_gbs,
                 "vs_value": (world * batch * sust_steps / (ms_s * 1e-3)) / value}

    # supplementary: the same K steps issued alternately on two streams (consecutive batches are
    # independent), which hides the start-up and drain of one launch behind the other -- what a
    # caller with several batches in flight gets; `value` stays the single-stream figure
    two = None
    if launches_per_step == 1:
        s2 = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for st2 in s2:
            st2.wait_event(t0)
        for i in range(args.steps):
            a, b, c = bufs[i % sets]
            plan.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, s2[i % 2].cuda_stream)
        for st2 in s2:
            ev = torch.cuda.Event()
            ev.record(st2)
            torch.cuda.current_stream().wait_event(ev)
        t1.record()
        torch.cuda.synchronize()
        ms2 = sh.max_over_ranks(t0.elapsed_time(t1), dev)
        two = {"value": world * batch * args.steps / (ms2 * 1e-3), "unit": "polymul/s", "streams": 2,
               "hbm_frac": None, "note": "same steps, alternating over two streams; not the headline value"}

    # e2e: host buffers (pinned) through the host-buffer API: H2D + kernel + D2H of every step inside the
    # timed region.  Headline: nttb200_polymul_batch_async with two products in flight (step i is queued,
    # then step i-1 is waited for: its c is complete in host memory) -- what a caller with a stream of
    # batches does; the plan's worker runs the queue as one stream of jobs, so the fill and drain of
    # consecutive products overlap.  `sync_value`: the same steps as back-to-back synchronous calls.
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    ha, hb, hc = mod.host_alloc((batch, n)), mod.host_alloc((batch, n)), mod.host_alloc((batch, n))
    hc2 = mod.host_alloc((batch, n))
    ha.array[:] = bufs[0][0].cpu().numpy()
    hb.array[:] = bufs[0][1].cpu().numpy()
    for _ in range(3):
        plan.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
    sh.barrier(local)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
    e2e_sync_s = sh.max_over_ranks(time.perf_counter() - t0, dev)
    ws = plan.wire_stats()                       # how the rows of one call crossed the link
    e2e_api = "nttb200_polymul_batch"
    e2e_s = e2e_sync_s
    hc_last = hc
    if hasattr(plan, "polymul_async_ptr") and os.environ.get("NTTB200_BENCH_E2E_SYNC", "0") != "1":
        outs = (hc, hc2)
        for i in range(2):                       # warm-up: worker thread, its first stream of jobs
            plan.polymul_async_ptr(outs[i].ptr, ha.ptr, hb.ptr, batch)
        plan.wait(0)
        hc.array[:1] = -1
        sh.barrier(local)
        t0 = time.perf_counter()
        prev = None
        for i in range(e2e_steps):
            t = plan.polymul_async_ptr(outs[i % 2].ptr, ha.ptr, hb.ptr, batch)
            if prev is not None:
                plan.wait(prev)
            prev = t
        plan.wait(prev)
        e2e_s = sh.max_over_ranks(time.perf_counter() - t0, dev)
        e2e_api = "nttb200_polymul_batch_async, two products in flight, nttb200_polymul_wait per step"
        if (e2e_steps - 1) % 2 == 1:             # the rows checked below are those of the LAST step
            hc_last = hc2
    hcd = torch.from_numpy(hc_last.array).to(dev)
    rows = np.unique(np.r_[0:9, batch - 4:batch, np.random.default_rng(11).integers(0, batch, 4096 if n <= 1024 else 24)])
    e2e_ok, _ = check_rows(mod, n, q, W.cyclic, bufs[0][0], bufs[0][1], hcd, rows)
    del hcd
    e2e_value = world * batch * e2e_steps / e2e_s
    e2e_sync_value = world * batch * e2e_steps / e2e_sync_s
    # what crossed the PCIe link in one call: 16-bit words for the rows the host pool
    # narrowed (half-word moduli, csrc/hostwire.c), the caller's 32-bit words for the rest
    if ws["rows16"] + ws["rows32"] == 0:
        ws["rows32"] = batch
    h2d_bytes = 2 * (2 * ws["rows16"] + 4 * ws["rows32"]) * n
    d2h_bytes = (2 * ws["result_rows16"] + 4 * (batch - ws["result_rows16"])) * n

    # the standalone NTT call surface (in place, 8n bytes per transform): HBM-bound kernels.
    # Run on the a-buffers of the rotating sets (after the product has been timed and checked).
    ntt_lines = {}
    try:
        for kind in ("mulntt_std2rev", "inttmul_rev2std_scaled") if "omega=" not in plan.describe() else ("ntt_std2rev", "intt_rev2std_scaled"):
            for i in range(3):
                plan.transform_dev(kind, bufs[i % sets][0].data_ptr(), batch, stream)
            torch.cuda.synchronize()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 30
            t0e.record()
            for i in range(reps):
                plan.transform_dev(kind, bufs[i % sets][0].data_ptr(), batch, stream)
            t1e.record()
            torch.cuda.synchronize()
            tms = t0e.elapsed_time(t1e) / reps
            gbs = 8 * n * batch / (tms * 1e-3) / 1e9
            ntt_lines[kind] = {"transforms_per_s": batch / (tms * 1e-3), "achieved_GBs": gbs, "frac_of_hbm_peak": gbs / peak_gbs}
    except Exception as ex:
        ntt_lines = {"error": repr(ex)}
    alg_bytes = W.alg_bytes
    # n <= 1024: one fused kernel per step.  n > 1024: the step is a pipeline of kernels; the
    # roofline is then stated for the whole pipeline (step time).
    roof_ms = per_launch_ms if launches_per_step == 1 else min(launch_ms, ms_local / args.steps)
    achieved = alg_bytes / (roof_ms * 1e-3) / 1e9
    logn = n.bit_length() - 1
    bflies = 3 * (n // 2) * logn
    modmuls = bflies + n                            # M
    plantard = W.plantard
    slots = slots_per_polymul(n, plantard, W.signed)
    rate = batch / (roof_ms * 1e-3)
    int_achieved = rate * 3 * modmuls
    slot_achieved = rate * slots
    traffic = TRAFFIC_NCU.get(args.workload)
    plan_desc, plan_psi = plan.describe(), plan.psi
    for h in (ha, hb, hc, hc2):
        h.free()
    W.close()

    # the other BASELINE configurations, measured by the same run (each a few hundred ms of device
    # time), and -- for N > 1 -- configs[2]/[3] as BASELINE.json states them: ONE fixed batch sliced
    # by nttb200_shard_bounds over the ranks (strong scaling)
    side, strong = {}, {}
    if args.workload == "c2" and not args.no_side_workloads:
        for name in ("c3", "c4", "c5"):
            try:
                side[name] = side_workload(mod, sh, torch, name, dev, rank, world, local, peak_gbs, imad_peak)
            except Exception as ex:
                side[name] = {"error": repr(ex)}
        for name in ("c3", "c4"):
            if world == 1:
                if "value" in side.get(name, {}):
                    strong[name] = {k: side[name][k] for k in ("value", "ms_per_step", "rows_per_gpu", "rows_total", "parity_ok")}
                continue
            try:
                full = 1 << WORKLOADS[name][3]
                lo, hi = mod.shard_bounds(full, world, rank)
                r = side_workload(mod, sh, torch, name, dev, rank, world, local, peak_gbs, imad_peak,
                                  rows_per_gpu=hi - lo, row_offset=lo)
                strong[name] = {k: r[k] for k in ("value", "ms_per_step", "rows_per_gpu", "rows_total", "parity_ok",
                                                  "steps", "clocks")}
            except Exception as ex:
                strong[name] = {"error": repr(ex)}

    line = {
        "metric": "polymul/s", "value": value, "unit": "polymul/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": desc, "n": n, "q": q, "batch_per_gpu": batch,
                   "l2": f"inputs larger than L2: {sets} rotating buffer sets ({sets * 3 * batch * row_bytes >> 20} MiB "
                         f"against 126 MB), no step finds its operands cached"},
        "setup": { "psi": plan_psi, "plan": plan_desc,
                   "inputs": "SURVEY 8d batch: splitmix64 stream, coefficient = next() % q, edge rows 0..7 "
                             "(0, q-1, delta_0, delta_{n-1}, KATs 1-4), the reference's coefficient files in row 8",
                   "l2": f"inputs rotate over {sets} buffer sets ({sets * 3 * batch * row_bytes >> 20} MiB) "
                         f"> 126 MB L2, so no step finds its operands cached",
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "host_binding": (f"rank 0 bound to {len(binding[0])} CPUs local to its GPU" if binding else "none")},
        "e2e": {"value": e2e_value, "unit": "polymul/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                "api": e2e_api + " (int32 host buffers in and out, pinned; stream ring; rows cross "
                       "PCIe as 16-bit words when the host thread pool narrows them, else as 32-bit words)",
                "sync_value": e2e_sync_value,
                "wire": {"rows_16bit": ws["rows16"], "rows_32bit": ws["rows32"],
                         "result_rows_16bit": ws["result_rows16"], "host_threads": ws["host_threads"],
                         "mode": os.environ.get("NTTB200_WIRE", "auto")},
                "host_bytes_per_step": 3 * batch * row_bytes, "parity_ok": e2e_ok},
        "gpu_launches": launches_per_step * args.steps,
        "rank_ms": rank_ms,       # device time of the K steps on every rank; `value` uses their maximum
        "clocks": clocks,
        "sustained": sustained,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "kernel": ((("polymul_splant_wide_kernel" if n in (512, 1024) else "polymul_splant_kernel") if W.signed
                                 else "polymul_plant_kernel") if plantard
                                else "polymul_small_kernel") if n <= 1024 else
                               "large-n product pipeline (whole step)",
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": roof_ms,
                     "traffic_source": TRAFFIC_SRC.get(args.workload)},
        "int_roofline": {"bound": "integer pipes / issue slots: the signed Plantard kernels level their instructions over "
                                  "the multiplier (fmaheavy) and ALU pipes and run at 82 - 83 % of the issue slots "
                                  "(ncu, profiles/); `frac` counts the multiplier slots the kernel spends, "
                                  "survey_3_per_modmul the 3 IMAD per modular multiplication of SURVEY 8d: read both "
                                  "next to roofline.frac, neither is the binding one",
                         "achieved": slot_achieved, "peak": imad_peak, "unit": "IMAD-slot lane-ops/s",
                         "frac": slot_achieved / imad_peak if imad_peak else None,
                         "slots_per_polymul": slots,
                         "survey_3_per_modmul": {"achieved": int_achieved, "imad_per_polymul": 3 * modmuls,
                                                 "frac": int_achieved / imad_peak if imad_peak else None},
                         "imad_hi_peak": imadhi_peak, "lazy_butterflies_per_s_peak": bfly_peak,
                         "note": "peak = independent IMAD chains on every SM, measured live "
                                 "(nttb200_measure_int_peak); IMAD.HI measured at half that rate so it "
                                 "counts 2 slots (slots_per_polymul in bench.py)",
                         "arith": ("plantard-signed" if W.signed else "plantard") if plantard else "shoup"},
        "parity_ok": parity_ok, "parity_rows_checked": parity_rows, "parity_checker": parity_checker,
        "two_streams": two,
        "standalone_ntt": ntt_lines,
        "workloads": side,
        "strong_scaling": strong,
    }
    if "arith=canon" in plan_desc:
        line["butterfly_roof"] = butterfly_roof_canon(mod, n, rate)
    if two:
        two["hbm_frac"] = two["value"] / world * 12 * n / 1e9 / peak_gbs
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(n, q, psi)
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    side_ok = all(r.get("parity_ok", False) for r in list(side.values()) + list(strong.values()) if "error" not in r)
    return 0 if (parity_ok and e2e_ok and side_ok) else 1


if __name__ == "__main__":
    sys.exit(main())
