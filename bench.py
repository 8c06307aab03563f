#!/usr/bin/env python
"""bench.py -- SymArrow n=32768 full eigen(): eigenpairs/s (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (restated: no Julia here)

One step = one full eigendecomposition of the synthetic ordered SymArrow (rng(421), D sorted
descending, arrow top last).  With N GPUs the eigenpair range 1..n is split into N contiguous blocks
(`k_partition`), every rank gets D,z,a and owns its block of U's columns; no collective is on the data path,
total work is fixed ("strong" scaling).  `value` is device-timed (CUDA events on the launching stream, max
over ranks) around the drop-in boundary call itself (ah_symarrow_eigen: host D, z of 0.5 MB validated, padded and
uploaded every step) with U left resident in HBM -- the figure with D, z already resident (ah_resolve) is reported
beside it in `config.value_resident_inputs`; `e2e` goes through the same C-ABI call with the eigenvectors copied
back into pinned host memory every step.  `--workload symdpr1|halfarrow` run BASELINE configs[2] / [3] (default
n = 16384) through the same harness; the default run appends the n = 65536 target as `target_n65536`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SymArrow n=%d full eigen(): eigenpairs/s"
FP64_INSTR_PER_TERM = 7.25    # DADD d, DFMA e, DMUL t, 3 DFMA reciprocal refinement, DFMA/DMUL accumulate + one DADD per tile sum of 4 (ah_device.cuh)


def gen_symarrow(n, seed=421):
    rng = np.random.default_rng(seed)
    N = n - 1
    D = np.sort(rng.random(N))[::-1].copy()
    z = rng.random(N)
    a = float(rng.random())
    return D, z, a


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU, sampled while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz, self.ok = index, [], 0, False, None, False
        self.interval = 0.01
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------ CPU arm
def time_oracle(n, D, z, a, sample_k, threads):
    """oracle (OpenMP static over k, one scratch arrow per thread = Threads.@threads + Bx[threadid()])."""
    import oracle
    lib = oracle.get_lib()
    tau = oracle.default_tau(n - 1)
    # a block of consecutive k around the middle of the spectrum, as in the reference's loop partition
    k0 = max(1, n // 2 - sample_k // 2)
    k1 = min(n, k0 + sample_k - 1)
    U = np.zeros((n, k1 - k0 + 1), order="F")
    t0 = time.perf_counter()
    import ctypes as C
    from oracle.lib import _dp, _p
    lam = np.zeros(k1 - k0 + 1)
    info = lib._info(k1 - k0 + 1)
    rc = lib.lib.ah_oracle_symarrow_eigen(n - 1, _p(D, _dp), _p(z, _dp), float(a), _p(tau, _dp), k0, k1, threads,
                                          _p(lam, _dp), _p(U, _dp), n, *lib._info_ptrs(info))
    assert rc == 0
    dt = time.perf_counter() - t0
    return (k1 - k0 + 1) / dt, dt, (k0, k1)


def cpu_sample_size(n, D, z, a, threads, target_s):
    rate, dt, _ = time_oracle(n, D, z, a, max(threads, 8), threads)      # pilot
    return int(max(threads, min(n, rate * target_s))), rate


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    D, z, a = gen_symarrow(n)
    threads = host_threads()
    sample, _ = cpu_sample_size(n, D, z, a, threads, 3.0)
    for _ in range(args.warmup):
        time_oracle(n, D, z, a, sample, threads)
    t0 = time.perf_counter()
    done = 0
    rng_k = None
    for _ in range(args.steps):
        _, _, rng_k = time_oracle(n, D, z, a, sample, threads)
        done += sample
    dt = time.perf_counter() - t0
    val = done / dt
    line = {
        "impl": "reference", "metric": METRIC % n, "value": val, "unit": "eigenpairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SymArrow n={n} full eigen(), GenSymArrow-like rng(421), ordered, arrow top last",
                   "note": "reference is Julia (not installable here): CPU restatement oracle/arrowhead_oracle.c, "
                           "OpenMP static over k = Threads.@threads; each step solves a bounded sample of k"},
        "cpu_baseline": {"value": val, "unit": "eigenpairs/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} consecutive eigenpairs k={rng_k[0]}..{rng_k[1]} of the n={n} problem per step"},
        "e2e": {"value": val, "unit": "eigenpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def gen_symdpr1(n, seed=421):
    rng = np.random.default_rng(seed)
    D = np.sort(rng.random(n))[::-1].copy()
    u = rng.random(n)
    rho = float(rng.random())
    return D, u, rho


def gen_halfarrow(nd, seed=421):
    rng = np.random.default_rng(seed)
    D = np.sort(np.abs(rng.random(nd) - 0.5))[::-1].copy()
    z = rng.random(nd + 1) - 0.5
    return D, z


class Problem:
    """one of the three matrix types of the hot path at order n (number of eigenpairs / singular triples = n)"""

    def __init__(self, kind, n):
        self.kind, self.n = kind, n
        if kind == "symarrow":
            self.D, self.z, self.a = gen_symarrow(n)
            self.npoles, self.ncols_v = n - 1, 0
            self.metric = METRIC % n
            self.workload = f"SymArrow n={n} full eigen(), GenSymArrow-like rng(421), ordered, arrow top last"
            self.entry = "ah_symarrow_eigen"
        elif kind == "symdpr1":
            self.D, self.z, self.a = gen_symdpr1(n)
            self.npoles, self.ncols_v = n, 0
            self.metric = f"SymDPR1 n={n} full eigen(): eigenpairs/s"
            self.workload = f"SymDPR1 n={n} full eigen(), GenSymDPR1-like rng(421), ordered, rho > 0"
            self.entry = "ah_symdpr1_eigen"
        else:
            self.D, self.z = gen_halfarrow(n - 1)
            self.a = None
            self.npoles, self.ncols_v = n - 1, n
            self.metric = f"HalfArrow length(z)={n} full svd(): singular triples/s"
            self.workload = f"HalfArrow nd={n - 1}, length(z)=nd+1={n} (SVD-update shape) full svd(), rng(421), ordered"
            self.entry = "ah_halfarrow_svd"

    def solve(self, ctx, k0, k1, U, V=None, host=False, D=None, z=None, into=None):
        D = self.D if D is None else D
        z = self.z if z is None else z
        kw = {("U_host" if host else "U_dev"): U, "ldu": self.n, "into": into}
        if self.kind == "symarrow":
            return ctx.symarrow(D, z, self.a, k0=k0, k1=k1, vectors=False, **kw)
        if self.kind == "symdpr1":
            return ctx.symdpr1(D, z, self.a, k0=k0, k1=k1, vectors=False, **kw)
        return ctx.halfarrow(D, z, k0=k0, k1=k1, vectors=False, U_dev=U, ldu=self.n, V_dev=V, ldv=self.n, into=into)

    def resolve(self, ctx, k0, k1, U, V=None, into=None):
        if self.kind == "halfarrow":
            return ctx.resolve(k0, k1, vectors=False, U_dev=U, ldu=self.n, V_dev=V, ldv=self.n, into=into)
        return ctx.resolve(k0, k1, vectors=False, U_dev=U, ldu=self.n, into=into)

    @property
    def vec_bytes(self):
        return self.n * 8 * (2 if self.kind == "halfarrow" else 1)


def fp64_denominator(ctx, sampler_index):
    """The FP64 roofline denominator, measured live and put on record: a dependent-free DFMA burst on every SM
    (ah_measure_fp64_peak, 8 independent chains x 512 threads x 4 CTAs per SM) with the SM clock sampled while it runs."""
    smp = ClockSampler(sampler_index)
    smp.interval = 0.002
    smp.start()
    best = 0.0
    for _ in range(6):
        g, ms = ctx.measure_fp64_peak()
        best = max(best, g)
    smp.stop_flag = True
    smp.join()
    term_gps, _ = ctx.measure_term_peak()
    clk = smp.summary()
    sm_count = 148
    nominal = 64.0 * sm_count * (clk["sm_max_mhz"] or 1965) * 1e6 / 1e9
    return {"fp64_ginstr_per_s": best, "fp64_tflops": 2.0 * best / 1e3, "sm_mhz_during_burst": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"],
            "nominal_ginstr_per_s": nominal, "frac_of_nominal": best / nominal,
            "how": "ah_measure_fp64_peak: dependent-free DFMA burst (8 chains/thread, 512 threads, 4 CTAs/SM, 2^15 iterations), best of 18, "
                   "CUDA events; nominal = 64 FP64 lanes x 148 SMs x max SM clock; MEASURED_PEAKS.json carries no FP64 entry",
            "term_burst_gterms_per_s": term_gps}


def timed_steps(torch, dist, dev, steps, fn, barrier):
    """exactly `steps` calls of fn between two CUDA events on the current stream; returns (max-over-ranks ms, last result, stats sums)"""
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    acc = {"launches": 0, "kernel_ms": 0.0, "main_kernel_ms": 0.0, "vec_ms": 0.0, "host_pre_ms": 0.0, "host_post_ms": 0.0}
    res = None
    for _ in range(steps):
        res = fn()
        for key in acc:
            acc[key] += res.stats[key]
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    barrier()
    t = torch.tensor([ms, float(acc["launches"])], dtype=torch.float64, device=dev)
    if dist is not None:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        return float(tm[0]), res, acc, int(ts[1])
    return ms, res, acc, acc["launches"]


def roofline_of(P, res, acc, steps, cnt, fp64, peaks, world):
    """k_bisect (round 0 + the Remedy-3 round): 7 FP64-pipe lane-instructions per secular term x steps_k x N / its device time"""
    fast = (res.status & 0xff) == 0            # (k_full's few eigenpairs are included: their first bisection ran in k_bisect)
    terms = float(res.steps[fast].astype(np.float64).sum()) * P.npoles
    t_main = acc["main_kernel_ms"] / steps * 1e-3
    instr_rate = FP64_INSTR_PER_TERM * terms / t_main / 1e9
    achieved_tf, peak_tf = 2.0 * instr_rate / 1e3, fp64["fp64_tflops"]
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_achieved = cnt * P.vec_bytes / (acc["vec_ms"] / steps * 1e-3) / 1e9
    traffic = ncu_pipe = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(f"{P.kind}_n{P.n}") or (tj if (P.kind == "symarrow" and P.n == tj.get("n")) else None)
        if ent and world == 1:
            traffic = ent["k_bisect_round0"]["dram_bytes_per_launch"]
            ncu_pipe = ent["k_bisect_round0"].get("sm__inst_executed_pipe_fp64")
    except Exception:
        pass
    return {
        "bound": "fp64", "kernel": f"k_bisect<{P.kind}> (round 0 + Remedy-3 round 1)", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf, "traffic": traffic,
        "peak_source": {"fp64_tflops": peak_tf, "measured": fp64, "ncu_sm__inst_executed_pipe_fp64_per_launch": ncu_pipe,
                        "ncu_file": "profiles/r2_*_full_*.md (same build; per-launch FP64-pipe instruction count = 7 x terms / 32 lanes + boundary code)"},
        "how": "EXECUTED FP64-pipe lane-instructions x2 (FMA-equivalent); 7.25 per secular term x sweeps_k x N terms per eigenpair (sweeps_k = the O(N) "
               "evaluations of the secular function k_bisect really ran, res.steps) / device time of "
               "k_bisect (round 0 + the Remedy-3 round, CUDA events on their streams inside the timed region); peak = DFMA burst measured "
               "live in this run (peak_source.measured)",
        "terms_per_launch": terms, "kernel_ms": t_main * 1e3, "fast_eigenpairs": int(res.stats["n_fast"]),
        "term_burst": {"gterms_per_s": fp64["term_burst_gterms_per_s"], "achieved_gterms_per_s": terms / t_main / 1e9,
                       "frac": terms / t_main / 1e9 / fp64["term_burst_gterms_per_s"],
                       "what": "secular terms/s of the same 7 FP64 + 1 MUFU.RCP64H mix from registers only (no tiles, no sweep "
                               "boundaries): what the ring, the barriers and the work distribution cost"},
        "hbm_store": {"kernel": "k_vec", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                      "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }


def numa_local_pinned(torch, local_rank, shape):
    """pinned host buffer whose pages are first touched by a thread bound to the GPU's NUMA node (cudaHostRegister of
    first-touched memory): eight ranks then do not all stream their D2H copies into the memory of one socket."""
    info = {"numa_node": None, "how": "torch pin_memory (no NUMA binding)"}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local_rank))
        node = None
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
            node = int(open(path).read().strip())
        except Exception:
            node = None
        if node is not None and node >= 0:
            cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            ids = []
            for part in cpus.split(","):
                lo, _, hi = part.partition("-")
                ids += list(range(int(lo), int(hi or lo) + 1))
            allowed = sorted(set(ids) & os.sched_getaffinity(0))
            if allowed:
                old = os.sched_getaffinity(0)
                os.sched_setaffinity(0, allowed)
                try:
                    t = torch.empty(shape, dtype=torch.float64)
                    t.zero_()                                  # first touch on the GPU's node
                    torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * 8, 0)
                finally:
                    os.sched_setaffinity(0, old)
                info = {"numa_node": node, "how": "first-touched by a thread bound to the GPU's NUMA node, then cudaHostRegister"}
                return t, info
    except Exception as e:                                     # fall through to the plain pinned allocation
        info["fallback_reason"] = f"{type(e).__name__}: {e}"[:120]
    return torch.empty(shape, dtype=torch.float64, pin_memory=True), info


def run_ours(args):
    # stdout carries exactly ONE JSON line (the driver parses it): everything else that libraries print there -- NCCL's
    # "NCCL version ..." banner at the first communicator, for one -- is sent to stderr until the line is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


def _run_ours(args):
    import torch
    import arrowhead_b200 as ab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ab.build_library()
    ctx = ab.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    n = args.n
    P = Problem(args.workload, n)
    k0, k1 = ab.k_partition(n, world, rank)
    cnt = k1 - k0 + 1
    # this rank's FULL-size U (n x n, column-major: row c of the tensor = column c): the rank fills its own column block in
    # place, so that the optional all-gather (ah_allgather_columns) needs no staging copy.  n = 32768: 8.6 GB per GPU.
    full_u = world > 1 and not args.no_allgather and n * n * 8 <= 40e9
    Ufull = torch.empty((n if full_u else cnt, n), dtype=torch.float64, device=dev)
    U = Ufull[k0 - 1:k1] if full_u else Ufull
    V = torch.empty((cnt, n), dtype=torch.float64, device=dev) if P.kind == "halfarrow" else None
    Uptr, Vptr = U.data_ptr(), (V.data_ptr() if V is not None else None)
    flush = torch.empty(20 << 20, dtype=torch.float64, device=dev)      # 160 MB > 126 MB L2 (float64: torch's 1-byte fill kernel is 10x slower)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from arrowhead_b200._lib import RangeResult
    outbuf = RangeResult(k0, k1)         # caller-owned per-k arrays, reused every step (as a Julia caller reuses Λ and κ)

    # L2 between steps: an explicit flush (a 160 MB write) only where the step itself does not already stream more than the
    # L2 through it.  Every step writes this rank's eigenvector block, cnt x n x 8 bytes (1.07 GB at n = 32768 on 8 GPUs, 8.6 GB on
    # one) -- nothing of the previous step survives that -- and the drop-in call uploads D, z afresh.
    need_flush = cnt * P.vec_bytes < 4 * (126 << 20)

    def step():                  # `value`: THE DROP-IN CALL -- host D, z in (validated, padded, uploaded every step), U resident
        if need_flush:
            flush.zero_()
        return P.solve(ctx, k0, k1, Uptr, Vptr, into=outbuf)

    def step_resident():         # beside it: inputs already resident in HBM (ah_resolve: launches + per-k results only)
        if need_flush:
            flush.zero_()
        return P.resolve(ctx, k0, k1, Uptr, Vptr, into=outbuf)

    fp64 = fp64_denominator(ctx, physical_gpu_index(local_rank))
    warm = max(args.warmup, 3)
    for _ in range(warm):
        res = step()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    if os.environ.get("BENCH_SAMPLER_INTERVAL"):
        sampler.interval = float(os.environ["BENCH_SAMPLER_INTERVAL"])
    sampler.start()
    ms_max, res, acc, launches_all = timed_steps(torch, dist, dev, args.steps, step, barrier)
    sampler.stop_flag = True
    value = n * args.steps / (ms_max * 1e-3)
    rsteps = max(3, min(args.steps, 10))
    ms_res, _, acc_res, _ = timed_steps(torch, dist, dev, rsteps, step_resident, barrier)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = roofline_of(P, res, acc, args.steps, cnt, fp64, peaks, world)
    # What the reference's algorithm would have executed: plain bisection (ARROWHEAD_B200_ACCEL=0) takes ~54 sweeps per
    # eigenpair where the certified skipping of k_bisect evaluates ~16 -- same final brackets, bit for bit.  One untimed
    # solve with the knob off counts them; `roofline.frac` stays the EXECUTED work over the FP64 peak.
    try:
        os.environ["ARROWHEAD_B200_ACCEL"] = "0"
        ctx_plain = ab.Context(local_rank)
        os.environ.pop("ARROWHEAD_B200_ACCEL", None)
        rp = P.solve(ctx_plain, k0, k1, Uptr, Vptr)
        torch.cuda.synchronize()
        fastp = (rp.status & 0xff) == 0
        plain_terms = float(rp.steps[fastp].astype(np.float64).sum()) * P.npoles
        same = bool(np.array_equal(rp.lam, res.lam, equal_nan=True) and np.array_equal(rp.status, res.status))
        t_main = roofline["kernel_ms"] * 1e-3
        roofline["reference_algorithm"] = {
            "terms_per_launch": plain_terms, "sweeps_per_eigenpair": float(rp.steps.mean()), "executed_sweeps_per_eigenpair": float(res.steps.mean()),
            "equivalent_tflops": 2.0 * FP64_INSTR_PER_TERM * plain_terms / t_main / 1e12,
            "equivalent_frac": 2.0 * FP64_INSTR_PER_TERM * plain_terms / t_main / 1e12 / roofline["peak"],
            "plain_bisection_kernel_ms": float(rp.stats["main_kernel_ms"]), "bitwise_identical_lambda_and_status": same,
            "what": "secular terms of the reference's plain bisection on the same input (one untimed solve with ARROWHEAD_B200_ACCEL=0) over "
                    "the accelerated kernel's time: > 1 means the certified skipping does more than the FP64 pipe could by brute force"}
        ctx_plain.close()
    except Exception as e:
        os.environ.pop("ARROWHEAD_B200_ACCEL", None)
        roofline["reference_algorithm"] = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- end to end: host D,z in, eigenvalues + Info + this rank's eigenvectors out to pinned host memory
    e2e = None
    if not args.no_e2e and P.kind != "halfarrow":
        Uh, numa = numa_local_pinned(torch, local_rank, (cnt, n))
        Dp = torch.from_numpy(P.D).pin_memory().numpy()
        zp = torch.from_numpy(P.z).pin_memory().numpy()
        e2e_steps = max(2, min(args.steps, 5))
        if world == 1 and P.kind == "symarrow":
            # the call a user makes: eigen(SymArrow) of the reference API (sort, deflation tests, solve, back-permutation),
            # eigenvectors streamed into a caller-owned pinned buffer
            A_user = ab.SymArrow(Dp, zp, P.a, n)
            Uh_np = Uh.numpy().T           # the same pinned memory seen as a Fortran-ordered n x n array

            class _R:                      # adapter: same fields as the k-range result used below
                pass

            def step_e2e():
                E, info = ab.eigen(A_user, ctx=ctx, out=Uh_np)
                r = _R(); r.lam = E.values; r.stats = info.stats
                return r
        else:
            def step_e2e():
                return P.solve(ctx, k0, k1, Uh.data_ptr(), host=True, D=Dp, z=zp)

        r2 = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r2 = step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine = float(r2.stats["d2h_bytes"]) * e2e_steps / dt / 1e9
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        rates = torch.zeros(world, dtype=torch.float64, device=dev); rates[rank] = mine
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(rates, op=dist.ReduceOp.SUM)
        barrier()
        hb = torch.tensor([float(r2.stats["h2d_bytes"]), float(r2.stats["d2h_bytes"])], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(hb, op=dist.ReduceOp.SUM)
        e2e = {"value": n * e2e_steps / float(tt[0]), "unit": "eigenpairs/s", "h2d_bytes_per_step": int(hb[0]),
               "d2h_bytes_per_step": int(hb[1]), "steps": e2e_steps,
               "d2h_gb_per_s_per_rank": [round(float(x), 2) for x in rates.cpu()], "pinned_buffer_rank0": numa,
               "what": ("eigen(SymArrow) of the Python mirror of the reference API (host sort + deflation tests + C-ABI solve + "
                        "permutation check), host D,z in, U/lambda/Info out to pinned host memory (wall clock)") if (world == 1 and P.kind == "symarrow") else
                       f"{P.entry} per rank with host D,z and this rank's U block/lambda/Info copied to pinned host memory (wall clock)"}
        # spot-check: the host copy equals the resident result
        assert np.array_equal(r2.lam, res.lam)
        assert torch.equal(Uh[cnt // 2].to(dev), U[cnt // 2])
        if "numa_node" in numa and numa["numa_node"] is not None:
            torch.cuda.cudart().cudaHostUnregister(Uh.data_ptr())
        del Uh

    # ---- all-gather of U over NVLink through the C ABI (north_star: only when the caller needs U on one device;
    # reported beside the headline, never inside it).  Every rank ends up with all n x n entries.
    allgather = None
    if dist is not None and full_u:
        uid = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.nccl_init(world, rank, uid[0])
        barrier()
        ctx.allgather_columns(Ufull.data_ptr(), n, n)                    # warm-up (NCCL channel setup)
        barrier()
        ms_ag = ctx.allgather_columns(Ufull.data_ptr(), n, n)
        tt = torch.tensor([ms_ag], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # every rank now holds every column: compare a few foreign columns with their owners' copies
        probe = torch.stack([Ufull[(r * n) // world + 3] for r in range(world)])
        ref = probe.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, probe))
        ok = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        recv_bytes = (n - cnt) * n * 8
        allgather = {"ms": float(tt[0]), "api": "ah_nccl_init + ah_allgather_columns (ncclAllGather in place, libnccl dlopen'ed by the library)",
                     "bytes_received_per_rank": recv_bytes, "gb_per_s_per_rank": recv_bytes / (float(tt[0]) * 1e-3) / 1e9,
                     "all_ranks_hold_identical_U": bool(ok.item() == 1.0)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and P.kind == "symarrow":
        threads = host_threads()
        sample, _ = cpu_sample_size(n, P.D, P.z, P.a, threads, 12.0)
        rate, dt, kr = time_oracle(n, P.D, P.z, P.a, sample, threads)
        cpu = {"value": rate, "unit": "eigenpairs/s", "cores": threads, "kind": "port",
               "sample": f"{sample} consecutive eigenpairs k={kr[0]}..{kr[1]} of the same n={n} problem, {dt:.1f} s, "
                         "oracle/arrowhead_oracle.c with OpenMP static over k"}

    # ---- the north-star target size as a sub-record of the default run (1 GPU, SymArrow): n = 65536, 3 steps
    target = None
    if world == 1 and P.kind == "symarrow" and n == 32768 and not args.no_target:
        try:
            del Ufull, U
            torch.cuda.empty_cache()
            nt = 65536
            Pt = Problem("symarrow", nt)
            Ut = torch.empty((nt, nt), dtype=torch.float64, device=dev)      # 34.4 GB resident
            def step_t():
                flush.zero_()
                return Pt.solve(ctx, 1, nt, Ut.data_ptr())
            for _ in range(2):
                rt = step_t()
            ms_t, rt, acc_t, _ = timed_steps(torch, None, dev, 3, step_t, barrier)
            rl = roofline_of(Pt, rt, acc_t, 3, nt, fp64, peaks, 1)
            lam = rt.lam
            target = {"workload": Pt.workload, "value": nt * 3 / (ms_t * 1e-3), "unit": "eigenpairs/s", "ms_per_step": ms_t / 3, "steps": 3,
                      "roofline_frac": rl["frac"], "k_bisect_ms": rl["kernel_ms"], "achieved_tflops": rl["achieved"],
                      "hbm_store_frac": rl["hbm_store"]["frac"],
                      "checks": {"strictly_descending": bool(np.all(np.diff(lam) < 0)),
                                 "interlacing": bool(np.all(lam[1:] < Pt.D) and np.all(lam[:-1] > Pt.D)),
                                 "trace_rel_err": float(abs(lam.sum() - (Pt.D.sum() + Pt.a)) / np.abs(lam).sum()),
                                 "status_all_ok": bool(((rt.status & 0xff) == 0).all())}}
            # plain bisection on the same problem (eigenvalues and Info only): bit-identical results, and what it would cost
            try:
                os.environ["ARROWHEAD_B200_ACCEL"] = "0"
                ctx_plain = ab.Context(local_rank)
                os.environ.pop("ARROWHEAD_B200_ACCEL", None)
                rp = Pt.solve(ctx_plain, 1, nt, None)
                torch.cuda.synchronize()
                target["plain_bisection"] = {"bitwise_identical_lambda_info_status": bool(
                                                 all(np.array_equal(getattr(rp, f), getattr(rt, f), equal_nan=True)
                                                     for f in ("lam", "sind", "kb", "kz", "knu", "krho", "qout", "status"))),
                                             "sweeps_per_eigenpair": float(rp.steps.mean()), "executed_sweeps_per_eigenpair": float(rt.steps.mean()),
                                             "k_bisect_ms": float(rp.stats["main_kernel_ms"])}
                ctx_plain.close()
            except Exception as e:
                os.environ.pop("ARROWHEAD_B200_ACCEL", None)
                target["plain_bisection"] = {"error": f"{type(e).__name__}: {e}"[:200]}
            del Ut
        except Exception as e:                                               # never lose the headline line to the sub-record
            target = {"error": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        line = {
            "metric": P.metric, "value": value, "unit": "eigenpairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": P.workload,
                       "parallelism": f"k-range x{world}", "U": "device-resident, column blocks per rank",
                       "timed_call": f"{P.entry} (the drop-in boundary call: host D, z validated, padded and uploaded every step; "
                                     "per-k results back to the host; eigenvectors stay in HBM)",
                       "value_resident_inputs": {"value": n * rsteps / (ms_res * 1e-3), "ms_per_step": ms_res / rsteps, "steps": rsteps,
                                                 "what": "the same step through ah_resolve (D, z already resident in HBM: no validation, padding, upload)"},
                       "l2": ("160 MB flush write before every step (inside the timed region)" if need_flush else
                              f"no explicit flush: every step streams this rank's eigenvector block of {cnt * P.vec_bytes / 1e9:.2f} GB "
                              f"(= {cnt * P.vec_bytes / (126 << 20):.1f} x the 126 MB L2) through the L2 and uploads D, z afresh"),
                       "allgather_U": allgather, "allgather_U_ms": allgather["ms"] if allgather else None,
                       "kernel_ms_per_step_rank0": acc["kernel_ms"] / args.steps,
                       "host_ms_per_step_rank0": {"before_first_launch": acc["host_pre_ms"] / args.steps, "after_sync": acc["host_post_ms"] / args.steps},
                       "kernel_ms_breakdown_rank0": {k: res.stats[k] for k in ("prep_ms", "bisect_ms", "rem3_ms", "vec_ms", "full_ms")},
                       "n_full_rank0": int(res.stats["n_full"]), "n_rem3_rank0": int(res.stats["n_rem3"]),
                       "total_steps_rank0": int(res.stats["total_steps"])},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches_all, "roofline": roofline,
            "cpu_baseline": cpu, "target_n65536": target,
        }
    else:
        line = None
    if dist is not None:
        dist.destroy_process_group()
    return line


def run_tdc(args):
    """Secondary workload (BASELINE.json configs[4]): tdc() of a SymTridiagonal, device-resident
    level-batched path.  Device-timed (CUDA events on the context's stream); the back-transform GEMMs (arrowhead7.jl:32-33)
    are the library's own DMMA kernel and get their own roofline line, measured shape by shape on the level shapes of this
    n and set beside cuBLAS (torch.bmm / torch.matmul on the same shapes, same box).  Not the headline metric."""
    import torch
    import arrowhead_b200 as ab
    n = args.n
    rng = np.random.default_rng(421)
    dv, ev = 2.0 + 1e-3 * rng.standard_normal(n), -1.0 + 1e-3 * rng.standard_normal(n - 1)
    ab.build_library()
    ctx = ab.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    warm = max(args.warmup, 1)
    for _ in range(warm):
        E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); E.vectors.free()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); E.vectors.free()
    ev1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / args.steps * 1e3
    ms = ev0.elapsed_time(ev1) / args.steps
    # ---- the GEMMs alone, on the level shapes of this decomposition: own kernel vs cuBLAS
    dmma_gfma, _ = ctx.measure_dmma_peak()
    fp64_ginstr, _ = ctx.measure_fp64_peak()
    shapes = {}
    for nb, m, nn, k in E.stats["gemm_shapes"]:
        shapes[(nb, m, nn, k)] = shapes.get((nb, m, nn, k), 0) + 1
    levels, t_own, t_blas, flops_all = [], 0.0, 0.0, 0.0
    for (nb, m, nn, k), reps in sorted(shapes.items(), key=lambda kv: -kv[0][1]):
        A = torch.randn((nb, k, m), dtype=torch.float64, device="cuda")      # [b][col][row] = column-major m x k
        B = torch.randn((nb, nn, k), dtype=torch.float64, device="cuda")     # column-major k x nn
        C = torch.empty((nb, nn, m), dtype=torch.float64, device="cuda")
        pa = [A[b].data_ptr() for b in range(nb)]; pb = [B[b].data_ptr() for b in range(nb)]; pc = [C[b].data_ptr() for b in range(nb)]
        def own():
            ctx.dgemm_batched(m, nn, k, pa, m, pb, k, pc, m)
        def blas():                                                          # row-major view: C^T = B^T A^T
            torch.bmm(B, A, out=C) if nb > 1 else torch.matmul(B[0], A[0], out=C[0])
        res = []
        for fn in (own, blas):
            fn(); torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(); fn(); a1.record(); torch.cuda.synchronize()
                best = min(best, a0.elapsed_time(a1))
            res.append(best)
        fl = 2.0 * nb * m * nn * k
        levels.append({"nb": nb, "m": m, "n": nn, "k": k, "calls": reps, "own_ms": res[0], "cublas_ms": res[1],
                       "own_tflops": fl / res[0] / 1e9, "cublas_tflops": fl / res[1] / 1e9})
        t_own += res[0] * reps; t_blas += res[1] * reps; flops_all += fl * reps
        del A, B, C
    peak_tf = 2.0 * dmma_gfma / 1e3
    roofline = {"bound": "tensor", "kernel": "k_dgemm_dmma (DMMA.8x8x4, csrc/ah_gemm.cuh)", "achieved": flops_all / t_own / 1e9, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": flops_all / t_own / 1e9 / peak_tf, "traffic": None,
                "peak_source": "DMMA.8x8x4 burst from registers measured live (ah_measure_dmma_peak); DFMA burst beside it: %.1f TFLOP/s" % (2 * fp64_ginstr / 1e3),
                "gemm_ms_per_decomposition": t_own, "cublas_ms_same_shapes": t_blas, "vs_cublas": t_blas / t_own, "levels": levels[:12]}
    cpu = None
    if not args.no_cpu:
        import oracle
        t0 = time.perf_counter()
        vals, _ = oracle.tdc(dv, ev, nthreads=host_threads())
        cpu_s = time.perf_counter() - t0
        cpu = {"value": cpu_s * 1e3, "unit": "ms", "cores": host_threads(), "kind": "port",
               "sample": f"one full tdc() n={n}: oracle drivers + OpenMP per-k solver, numpy/OpenBLAS back-transform",
               "max_abs_eigenvalue_diff": float(np.max(np.abs(np.sort(vals) - np.sort(E.values))))}
    st = {k: v for k, v in E.stats.items() if k != "gemm_shapes"}
    print(json.dumps({"metric": f"SymTridiagonal n={n} tdc(): ms per decomposition (eigenvectors resident on the GPU)",
                      "value": ms, "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": warm,
                      "higher_is_better": False, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"tdc n={n}: 2,-1 Toeplitz + 1e-3 noise, rng(421)", "stats": st,
                                 "timing": "CUDA events on the context's stream around the whole decomposition", "host_wall_ms": wall_ms},
                      "roofline": roofline, "cpu_baseline": cpu}), flush=True)


DEFAULT_N = {"symarrow": 32768, "symdpr1": 16384, "halfarrow": 16384, "tdc": 8192}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=None,
                    help="matrix order (under torchrun spell it --size: its own parser treats --n as an abbreviation)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="symarrow", choices=["symarrow", "symdpr1", "halfarrow", "tdc"])
    ap.add_argument("--no-allgather", action="store_true", help="N>1: skip the NCCL all-gather of U (timed after the headline, reported in config)")
    ap.add_argument("--allgather", action="store_true", help="(kept for compatibility: the all-gather now always runs at N>1)")
    ap.add_argument("--no-target", action="store_true", help="skip the n=65536 sub-record of the default 1-GPU run")
    args = ap.parse_args()
    if args.n is None:
        args.n = DEFAULT_N[args.workload]
    if args.workload == "tdc":
        run_tdc(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
