#!/usr/bin/env python3
"""bench.py -- polymul/s of the batched NTT polynomial multiplier (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

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
TRAFFIC_NCU = {"c2": 134264064 + 35679488}
TRAFFIC_SRC = {"c2": "profiles/r1_c2_polymul_plant_n256_v5_ncu_full.txt: dram__bytes_read.sum 134.26 MB + "
                     "dram__bytes_write.sum 35.68 MB per launch (algorithmic 201.3 MB; part of c stays dirty in L2)"}
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


def run_reference_arm(args, n, q, psi, logb, desc):
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
        "config": {"workload": desc, "n": n, "q": q, "batch_per_step": sample,
                   "note": "CPU arm: host cores only, no GPU, no host<->device copies"},
        "cpu_baseline": {"value": value, "unit": "polymul/s", "cores": cores, "kind": kind,
                         "sample": f"{what}; each step = {sample} polymuls ({reps} passes over {per} rows "
                                   f"per process, {cores} processes = one per core; the workload has {full}), "
                                   f"wall clock including the operand restore the reference needs"},
        "e2e": {"value": value, "unit": "polymul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    args = ap.parse_args()
    n, q, psi, logb, desc = WORKLOADS[args.workload]
    batch = 1 << logb

    if args.impl == "reference":
        return run_reference_arm(args, n, q, psi, logb, desc)

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
    # narrow / widen the rows of the host-buffer call (csrc/hostwire.c)
    os.environ.setdefault("NTTB200_HOST_THREADS", str(max(1, host_cores() // max(1, world))))
    warmup = max(3, args.warmup)

    cyclic = (q - 1) % (2 * n) != 0                  # q=3329, n=256: no 512-th root of unity
    plan = mod.Plan(n, q, psi, cyclic=cyclic)
    row_bytes = n * 4
    # distinct buffer sets so that successive steps never find their inputs in the 126 MB L2
    sets = max(2, -(-3 * L2_BYTES // (3 * batch * row_bytes)) + 1)
    g = torch.Generator(device=dev).manual_seed(SEED % (2**31) + rank)
    bufs = []
    for _ in range(sets):
        a = torch.randint(0, q, (batch, n), dtype=torch.int32, device=dev, generator=g)
        b = torch.randint(0, q, (batch, n), dtype=torch.int32, device=dev, generator=g)
        c = torch.empty_like(a)
        bufs.append((a, b, c))
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        a, b, c = bufs[i % sets]
        plan.polymul_dev(c.data_ptr(), a.data_ptr(), b.data_ptr(), batch, stream)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(warmup):
        step(i)
    launches_per_step = mod.last_launch_count()
    torch.cuda.synchronize()
    sh.barrier(local)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.timed(True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    sampler.timed(False)
    sh.barrier(local)
    ms_local = e0.elapsed_time(e1)
    ms = sh.max_over_ranks(ms_local, dev)
    clocks = sampler.stop()
    value = world * batch * args.steps / (ms * 1e-3)

    # per-launch duration of the dominant kernel, CUDA events on the launching stream
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
    for i, (s, e) in enumerate(evs):
        s.record()
        step(i)
        e.record()
    torch.cuda.synchronize()
    launch_ms = statistics.mean(s.elapsed_time(e) for s, e in evs)
    # a back-to-back stream hides launch gaps: the K-step region gives the better per-launch figure
    per_launch_ms = min(launch_ms, ms_local / args.steps) / max(1, launches_per_step)

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

    # parity spot check of what was just timed (never skip work silently)
    from oracle import loader
    O = loader.Oracle()
    a, b, c = bufs[(args.steps - 1) % sets]
    idx = torch.tensor([0, 1, batch // 2, batch - 1], device=dev)
    variant = loader.PRODUCT_CYCLIC if cyclic else loader.PRODUCT_MERGED
    want = O.product(n, q, a[idx].cpu().numpy(), b[idx].cpu().numpy(), variant)
    parity_ok = bool((c[idx].cpu().numpy() == want).all())

    # e2e: host buffers (pinned) through nttb200_polymul_batch: H2D + kernel + D2H per step
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    ha, hb, hc = mod.host_alloc((batch, n)), mod.host_alloc((batch, n)), mod.host_alloc((batch, n))
    ha.array[:] = bufs[0][0].cpu().numpy()
    hb.array[:] = bufs[0][1].cpu().numpy()
    for _ in range(3):
        plan.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
    sh.barrier(local)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.polymul_host_ptr(hc.ptr, ha.ptr, hb.ptr, batch)
    e2e_s = sh.max_over_ranks(time.perf_counter() - t0, dev)
    rows = np.unique(np.r_[0:2, batch // 2 - 1:batch // 2 + 1, batch - 2:batch,
                           np.random.default_rng(7).integers(0, batch, 26)])
    e2e_ok = bool((hc.array[rows] == O.product(n, q, ha.array[rows], hb.array[rows], variant)).all())
    e2e_value = world * batch * e2e_steps / e2e_s
    # what crossed the PCIe link in the last timed call: 16-bit words for the rows the host pool
    # narrowed (half-word moduli, csrc/hostwire.c), the caller's 32-bit words for the rest
    ws = plan.wire_stats()
    if ws["rows16"] + ws["rows32"] == 0:
        ws["rows32"] = batch
    h2d_bytes = 2 * (2 * ws["rows16"] + 4 * ws["rows32"]) * n
    d2h_bytes = (2 * ws["result_rows16"] + 4 * (batch - ws["result_rows16"])) * n

    peak_gbs, peak_src = measured_peaks()
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
    alg_bytes = 12 * n * batch                      # read a, read b, write c (int32 API)
    # n <= 1024: one fused kernel per step.  n > 1024: the step is a pipeline of 3 kernels per
    # L2-resident batch chunk; the roofline is then stated for the whole pipeline (step time).
    roof_ms = per_launch_ms if launches_per_step == 1 else min(launch_ms, ms_local / args.steps)
    achieved = alg_bytes / (roof_ms * 1e-3) / 1e9
    logn = n.bit_length() - 1
    bflies = 3 * (n // 2) * logn
    modmuls = bflies + n                            # M
    # fmaheavy issue slots per product as the kernels are written: Shoup butterfly = IMAD.HI
    # (2 slots: measured half rate) + 2 IMAD; pointwise Montgomery = 2 IMAD.HI + 2 IMAD; the
    # n^-1 scaling costs one extra Shoup multiplication on the sum branch of the last stage
    plantard = "plantard" in plan.describe()
    if plantard and n <= 256:   # half-word second product: 2 IMAD (+ 2 shifts on the ALU pipe) per butterfly
        slots = 2 * bflies + 4 * n + 2 * (n // 2)
    elif plantard:    # IMAD + IMAD.HI per butterfly (3), 2 IMAD + IMAD.HI pointwise (4), 3 per n^-1 scale
        slots = 3 * bflies + 4 * n + 3 * (n // 2)
    else:
        slots = 4 * bflies + 6 * n + 4 * (n // 2)
    imad_peak = mod.measure_int_peak(0)
    imadhi_peak = mod.measure_int_peak(1)
    bfly_peak = mod.measure_int_peak(3)
    rate = batch / (roof_ms * 1e-3)
    int_achieved = rate * 3 * modmuls
    slot_achieved = rate * slots
    traffic = TRAFFIC_NCU.get(args.workload)

    line = {
        "metric": "polymul/s", "value": value, "unit": "polymul/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": desc, "n": n, "q": q, "psi": plan.psi, "batch_per_gpu": batch,
                   "plan": plan.describe(),
                   "l2": f"inputs rotate over {sets} buffer sets ({sets * 3 * batch * row_bytes >> 20} MiB) "
                         f"> 126 MB L2, so no step finds its operands cached",
                   "parallelism": f"batch-sharded x{world}, no collective"},
        "e2e": {"value": e2e_value, "unit": "polymul/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                "api": "nttb200_polymul_batch (int32 host buffers in and out, pinned; stream ring; rows cross "
                       "PCIe as 16-bit words when the host thread pool narrows them, else as 32-bit words)",
                "wire": {"rows_16bit": ws["rows16"], "rows_32bit": ws["rows32"],
                         "result_rows_16bit": ws["result_rows16"], "host_threads": ws["host_threads"],
                         "mode": os.environ.get("NTTB200_WIRE", "auto")},
                "host_bytes_per_step": 3 * batch * row_bytes, "parity_ok": e2e_ok},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "kernel": ("polymul_plant_kernel" if plantard else "polymul_small_kernel") if n <= 1024 else
                               "large_cols_fwd + large_rows_polymul + large_cols_inv (whole step)",
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": roof_ms,
                     "traffic_source": TRAFFIC_SRC.get(args.workload)},
        "int_roofline": {"bound": "multiplier (fmaheavy) pipe; the n<=256 Plantard kernel trades multiplier slots for "
                                  "ALU instructions, so issue slots and the two integer pipes are near level "
                                  "(ncu, profiles/): read this fraction next to roofline.frac, not as the binding one",
                         "achieved": slot_achieved, "peak": imad_peak, "unit": "IMAD-slot lane-ops/s",
                         "frac": slot_achieved / imad_peak if imad_peak else None,
                         "slots_per_polymul": slots,
                         "survey_3_per_modmul": {"achieved": int_achieved, "imad_per_polymul": 3 * modmuls,
                                                 "frac": int_achieved / imad_peak if imad_peak else None},
                         "imad_hi_peak": imadhi_peak, "lazy_butterflies_per_s_peak": bfly_peak,
                         "note": "peak = independent IMAD chains on every SM, measured live "
                                 "(nttb200_measure_int_peak); IMAD.HI measured at half that rate so it "
                                 "counts 2 slots. Shoup kernels: butterfly 4, pointwise 6, n^-1 scale 4 per "
                                 "pair; Plantard kernel (q<=12385): butterfly 3, pointwise 4, scale 3 at n=512/1024; at n<=256 butterfly and scale 2 (IMAD, SHF, IMAD, SHF: the second product takes only the upper half of the first)",
                         "arith": "plantard" if plantard else "shoup"},
        "parity_ok": parity_ok,
        "two_streams": two,
        "standalone_ntt": ntt_lines,
    }
    if two:
        two["hbm_frac"] = two["value"] / world * 12 * n / 1e9 / peak_gbs
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(n, q, psi)
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    for h in (ha, hb, hc):
        h.free()
    plan.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0 if (parity_ok and e2e_ok) else 1


if __name__ == "__main__":
    sys.exit(main())
