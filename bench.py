#!/usr/bin/env python
"""bench.py -- SymArrow n=32768 full eigen(): eigenpairs/s (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (restated: no Julia here)

One step = one full eigendecomposition of the synthetic ordered SymArrow (rng(421), D sorted
descending, arrow top last).  With N GPUs the eigenpair range 1..n is split into N contiguous blocks
(`k_partition`), every rank gets D,z,a and owns its block of U's columns; no collective is on the data path,
total work is fixed ("strong" scaling).  `value` is device-timed (CUDA events on the launching stream, max
over ranks) with inputs passed as host arrays of 0.5 MB and U left resident in HBM; `e2e` goes through the
same C-ABI call with the eigenvectors copied back into pinned host memory every step.
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
FP64_INSTR_PER_TERM = 7.0     # DADD d, DFMA e, DMUL t, 3 DFMA reciprocal refinement, DFMA accumulate (ah_device.cuh)


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
def run_ours(args):
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
    D, z, a = gen_symarrow(n)
    k0, k1 = ab.k_partition(n, world, rank)
    cnt = k1 - k0 + 1
    U = torch.empty((cnt, n), dtype=torch.float64, device=dev)          # this rank's columns of U (column-major block)
    flush = torch.empty(32 << 20, dtype=torch.float64, device=dev)      # 256 MB > 126 MB L2 (float64: torch's 1-byte fill kernel is 10x slower than the memory)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def upload_and_step():       # the drop-in call: validates, pads and uploads D, z, then solves
        flush.zero_()
        return ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)

    def step():                  # `value`: inputs already resident in HBM (ah_resolve); `e2e` below uploads them every step
        flush.zero_()
        return ctx.resolve(k0, k1, vectors=False, U_dev=U.data_ptr(), ldu=n)

    fp64_ginstr, _ = ctx.measure_fp64_peak()
    term_gps, _ = ctx.measure_term_peak()      # the inner block's instruction mix from registers only
    res = upload_and_step()
    for _ in range(max(args.warmup, 3)):
        res = step()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    launches = 0
    kern_ms = main_ms = vec_ms = 0.0
    for _ in range(args.steps):
        res = step()
        launches += res.stats["launches"]
        kern_ms += res.stats["kernel_ms"]
        main_ms += res.stats["main_kernel_ms"]
        vec_ms += res.stats["vec_ms"]
    ev1.record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    barrier()
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms_max, launches_all = float(tm[0]), int(ts[1])
    else:
        ms_max, launches_all = ms, launches
    value = n * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (k_bisect) on this rank, from the same timed region
    N = n - 1
    fast = (res.status & 0xff) == 0            # (k_full's few eigenpairs are included: their first bisection ran in k_bisect)
    n_fast = res.stats["n_fast"]
    terms_fast = float(res.steps[fast].astype(np.float64).sum()) * N   # k_bisect (both rounds): one sweep of N secular terms per bisection step
    t_main = main_ms / args.steps * 1e-3
    instr_rate = FP64_INSTR_PER_TERM * terms_fast / t_main / 1e9       # G lane-instructions/s
    achieved_tf = 2.0 * instr_rate / 1e3
    peak_tf = 2.0 * fp64_ginstr / 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_achieved = cnt * n * 8 / (vec_ms / args.steps * 1e-3) / 1e9    # U is written by k_vec (its span also hosts the Remedy-3 round)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["k_bisect_round0"]["dram_bytes_per_launch"] if n == tj.get("n") and world == 1 else None
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": "k_bisect<SymArrow> (round 0 + Remedy-3 round 1)", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf, "traffic": traffic,
        "how": "FP64-pipe lane-instructions x2 (FMA-equivalent); 7 per secular term x steps_k x N terms per "
               "eigenpair / device time of k_bisect (round 0 + the Remedy-3 round, CUDA events on their streams inside "
               "the timed region); peak = dependent-free DFMA burst measured live (MEASURED_PEAKS.json has no FP64 entry)",
        "terms_per_launch": terms_fast, "kernel_ms": t_main * 1e3, "fast_eigenpairs": int(n_fast),
        "term_burst": {"gterms_per_s": term_gps, "achieved_gterms_per_s": terms_fast / t_main / 1e9,
                       "frac": terms_fast / t_main / 1e9 / term_gps,
                       "what": "secular terms/s of the same 7 FP64 + 1 MUFU.RCP64H mix from registers only (no tiles, no sweep "
                               "boundaries): what the ring, the barriers and the work distribution cost"},
        "hbm_store": {"kernel": "k_vec", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                      "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- end to end: host D,z in, eigenvalues + Info + this rank's eigenvectors out to pinned host memory
    e2e = None
    if not args.no_e2e:
        Uh = torch.empty((cnt, n), dtype=torch.float64, pin_memory=True)
        Dp = torch.from_numpy(D).pin_memory().numpy()
        zp = torch.from_numpy(z).pin_memory().numpy()
        e2e_steps = max(2, min(args.steps, 5))

        if world == 1:
            # the call a user makes: eigen(SymArrow) of the reference API (sort, deflation tests, solve, back-permutation),
            # eigenvectors streamed into a caller-owned pinned buffer
            A_user = ab.SymArrow(Dp, zp, a, n)
            Uh_np = Uh.numpy().T           # the same pinned memory seen as a Fortran-ordered n x n array

            class _R:                      # adapter: same fields as the k-range result used below
                pass

            def step_e2e():
                E, info = ab.eigen(A_user, ctx=ctx, out=Uh_np)
                r = _R(); r.lam = E.values; r.stats = info.stats
                return r
        else:
            def step_e2e():
                return ctx.symarrow(Dp, zp, a, k0=k0, k1=k1, vectors=False, U_host=Uh.data_ptr(), ldu=n)

        r2 = step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r2 = step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        barrier()
        hb = torch.tensor([float(r2.stats["h2d_bytes"]), float(r2.stats["d2h_bytes"])], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(hb, op=dist.ReduceOp.SUM)
        e2e = {"value": n * e2e_steps / float(tt[0]), "unit": "eigenpairs/s", "h2d_bytes_per_step": int(hb[0]),
               "d2h_bytes_per_step": int(hb[1]), "steps": e2e_steps,
               "what": ("eigen(SymArrow) of the Python mirror of the reference API (host sort + deflation tests + C-ABI solve + "
                        "permutation check), host D,z in, U/lambda/Info out to pinned host memory (wall clock)") if world == 1 else
                       "ah_symarrow_eigen per rank with host D,z and this rank's U block/lambda/Info copied to pinned host memory (wall clock)"}
        # spot-check: the host copy equals the resident result
        assert np.array_equal(r2.lam, res.lam)
        assert torch.equal(Uh[cnt // 2].to(dev), U[cnt // 2])

    allgather_ms = None
    if dist is not None and args.allgather:
        # optional, outside the headline: every rank ends up with all of U (NCCL all_gather of the column blocks)
        from arrowhead_b200.sharding import allgather_columns
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        Ufull = allgather_columns(U, n, n, world, rank, dist)
        a1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        allgather_ms = float(tt[0])
        assert Ufull.shape == (n, n) and torch.equal(Ufull[k0 - 1:k1], U)
        del Ufull

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        sample, _ = cpu_sample_size(n, D, z, a, threads, 12.0)
        rate, dt, kr = time_oracle(n, D, z, a, sample, threads)
        cpu = {"value": rate, "unit": "eigenpairs/s", "cores": threads, "kind": "port",
               "sample": f"{sample} consecutive eigenpairs k={kr[0]}..{kr[1]} of the same n={n} problem, {dt:.1f} s, "
                         "oracle/arrowhead_oracle.c with OpenMP static over k"}

    if rank == 0:
        line = {
            "metric": METRIC % n, "value": value, "unit": "eigenpairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"SymArrow n={n} full eigen(), GenSymArrow-like rng(421), ordered, arrow top last",
                       "parallelism": f"k-range x{world}", "U": "device-resident, column blocks per rank",
                       "inputs": "D, z (0.5 MB) resident in HBM, uploaded once before the timed region (ah_resolve); e2e uploads them every step",
                       "l2": "256 MB flush write before every step; the step itself streams out n^2*8 bytes >> L2",
                       "allgather_U_ms": allgather_ms,
                       "kernel_ms_per_step_rank0": kern_ms / args.steps,
                       "kernel_ms_breakdown_rank0": {k: res.stats[k] for k in ("prep_ms", "bisect_ms", "rem3_ms", "vec_ms", "full_ms")},
                       "n_full_rank0": int(res.stats["n_full"]), "n_rem3_rank0": int(res.stats["n_rem3"]),
                       "total_steps_rank0": int(res.stats["total_steps"])},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches_all, "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_tdc(args):
    """Secondary workload (BASELINE.json configs[4]): tdc() of a SymTridiagonal, device-resident
    level-batched path vs the oracle's restatement of arrowhead7.jl:10-36 on the host cores.  Not the headline metric."""
    import arrowhead_b200 as ab
    n = args.n if args.n != 32768 else 8192
    rng = np.random.default_rng(421)
    dv, ev = 2.0 + 1e-3 * rng.standard_normal(n), -1.0 + 1e-3 * rng.standard_normal(n - 1)
    ab.build_library()
    ctx = ab.Context(0)
    for _ in range(max(args.warmup, 1)):
        E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); E.vectors.free()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        E = ab.tdc_device(dv, ev, ctx=ctx, vectors="device"); E.vectors.free()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    cpu = None
    if not args.no_cpu:
        import oracle
        t0 = time.perf_counter()
        vals, _ = oracle.tdc(dv, ev, nthreads=host_threads())
        cpu_s = time.perf_counter() - t0
        cpu = {"value": cpu_s * 1e3, "unit": "ms", "cores": host_threads(), "kind": "port",
               "sample": f"one full tdc() n={n}: oracle drivers + OpenMP per-k solver, numpy/OpenBLAS back-transform",
               "max_abs_eigenvalue_diff": float(np.max(np.abs(np.sort(vals) - np.sort(E.values))))}
    print(json.dumps({"metric": f"SymTridiagonal n={n} tdc(): ms per decomposition (eigenvectors resident on the GPU)",
                      "value": ms, "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1),
                      "higher_is_better": False, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"tdc n={n}: 2,-1 Toeplitz + 1e-3 noise, rng(421)", "stats": E.stats,
                                 "timing": "host wall clock around tdc_device (host-orchestrated levels)"},
                      "cpu_baseline": cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=32768,
                    help="matrix order (under torchrun spell it --size: its own parser treats --n as an abbreviation)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="symarrow", choices=["symarrow", "tdc"])
    ap.add_argument("--allgather", action="store_true", help="N>1: also time an NCCL all-gather of U (reported separately)")
    args = ap.parse_args()
    if args.workload == "tdc":
        run_tdc(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
