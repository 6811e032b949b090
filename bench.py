#!/usr/bin/env python
"""bench.py — kriged grid points / second of the B200 backend='cuda' execute() path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): OrdinaryKriging 2-D, N=5000 random-scatter
data (seed 1002), 1000x1000 grid, exponential variogram [1.0, 300, 0.05], fp64.
One step = one full execute(): assemble + factor + krige every grid point (the reference re-assembles
and re-inverts on every call, ok.py:847,663 — so does every timed step here; nothing is cached).

  value : whole-step throughput with the point generation on device and outputs left in HBM
          (kb200_set_problem + kb200_execute_grid_dev); timed with CUDA events on the launching stream.
  e2e   : the same step through the public class API (OrdinaryKriging.execute('grid', ...,
          backend='cuda')) with host buffers: H2D of data/axes and D2H of (z, sigma^2) inside the region.
  roofline : the fused solve kernel (solve_kernel_f64) against the MEASURED fp64 GEMM rate of this GPU
          (torch.matmul 8192^3 taken in this run; MEASURED_PEAKS.json holds no fp64 entry), using the
          algorithmic 2*(N+1)^2 flop per grid point of SURVEY.md §8(d).
  cpu_baseline : the oracle port of the reference's inverse x RHS path (oracle/krige_oracle.py) on the
          box's host cores, on a bounded sample of the same grid.

N>1 (torchrun, one rank per GPU): rank 0 factors, one NCCL broadcast ships the factor blob, every rank
kriges a contiguous slice of a grid that grows with the GPU count (weak scaling: 1000 x 1000*N points);
the same run also times the fixed 1000x1000 grid split N ways and reports it under config.strong.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "kriged grid points/sec (N data -> M grid)"
UNIT = "points/s"
N_DATA = 5000
GRID = 1000
PARAMS = [1.0, 300.0, 0.05]
MODEL = "exponential"
SEED = 1002


def workload():
    import cases
    xyz, val = cases.synth_data(SEED, N_DATA, 2)
    gx = np.linspace(0.0, 1000.0, GRID)
    gy = np.linspace(0.0, 1000.0, GRID)
    return xyz, val, gx, gy


def config_dict(n_gpus, extra=None):
    c = {
        "workload": "OrdinaryKriging 2D, N=5000 data, 1000x1000 grid, exponential variogram, fp64 (BASELINE configs[1])",
        "n_data": N_DATA, "grid": [GRID, GRID], "variogram": MODEL, "variogram_parameters": PARAMS,
        "seed": SEED, "l2": "flush: a 256 MiB buffer is rewritten between timed steps; each step also rewrites "
                            "the 3x210 MB factor workspaces",
        "parallelism": "grid-point sharding, %d GPU(s), one NCCL broadcast of the factor blob" % n_gpus,
    }
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "samples": len(s), "reasons": sorted(self.reasons)}


def measure_fp64_peak(torch):
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n**3 / (best * 1e-3) / 1e12


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU arithmetic for this path on the host cores: the oracle
    port of backend='vectorized' (_get_kriging_matrix + scipy.linalg.inv + inverse x RHS, ok.py:626-683)
    and, when oracle/_ref is built, the reference's compiled backend='C' twin (cok.pyx:_c_exec_loop); the
    faster of the two is the reported value.
    Each step = a bounded slab of the same 1000x1000 grid; the matrix inverse is memoised outside
    the timed region (set-up reported separately), which favours the CPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import krige_oracle as ko
    import scipy.linalg
    from scipy.spatial.distance import cdist
    xyz, val, gx, gy = workload()
    stored = ko.stored_parameters(MODEL, PARAMS)
    t0 = time.perf_counter()
    a = ko.kriging_matrix(xyz, MODEL, stored)
    a_inv = scipy.linalg.inv(a)
    setup_s = time.perf_counter() - t0
    slab = 10000                       # 10 rows of the grid per step

    def step(i):
        rows = np.arange(10) + 10 * (i % 100)
        G = ko.grid_points([gx, gy[rows]])
        bd = cdist(G, xyz)
        b = np.ones((G.shape[0], N_DATA + 1))
        b[:, :N_DATA] = -ko.variogram(MODEL, stored, bd)
        b[:, :N_DATA][np.absolute(bd) <= ko.EPS] = 0.0
        x = a_inv @ b.T
        z = x[:N_DATA, :].T @ val
        ss = -np.einsum("ij,ji->i", b, x)
        return z, ss

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = slab / dt
    cores = os.cpu_count()
    kind = "port"
    sample = ("10 grid rows (10000 points) per step; A^-1 memoised outside the timed "
              "region (set-up %.2f s: matrix + scipy.linalg.inv)" % setup_s)
    other = None
    try:
        # the reference's own compiled twin of the path (lib/cok.pyx:_c_exec_loop, backend='C'), when
        # oracle/_ref was built: per-point dgemv over the inverse. Two sample sizes separate its internal
        # set-up (scipy.linalg.inv inside the call) from the per-point rate.
        from oracle import ref_native as rn
        if rn.available():
            G = ko.grid_points([gx, gy[:1]])
            t1 = time.perf_counter(); rn.exec_loop(xyz, G[:100], val, MODEL, stored); t1 = time.perf_counter() - t1
            t2 = time.perf_counter(); rn.exec_loop(xyz, G[:600], val, MODEL, stored); t2 = time.perf_counter() - t2
            nat = 500.0 / max(1e-9, t2 - t1)
            other = {"value": nat, "unit": UNIT, "kind": "reference",
                     "sample": "oracle/_ref cok._c_exec_loop (the reference's backend='C'): 600 vs 100 grid points, "
                               "steady per-point rate (its internal matrix inverse excluded)"}
            if nat > value:          # report the FASTER CPU implementation of the path as the reference arm
                other, value, kind, sample = (
                    {"value": value, "unit": UNIT, "kind": "port", "sample": sample}, nat, "reference", other["sample"])
                dt = slab / value
    except Exception as e:  # noqa: BLE001
        other = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, {"sample": "10000-point slabs of the grid per step"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "other_cpu_implementation": other},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_sample():
    """Bounded CPU baseline in the default run (rank 0, N=1): ~10-20 s of host work."""
    from oracle import krige_oracle as ko
    xyz, val, gx, gy = workload()
    stored = ko.stored_parameters(MODEL, PARAMS)
    G = ko.grid_points([gx, gy[:10]])          # 10000 points
    t0 = time.perf_counter()
    ko.krige_chunked(xyz, val, MODEL, stored, G, chunk=10000)
    t_all = time.perf_counter() - t0
    import scipy.linalg
    t1 = time.perf_counter()
    scipy.linalg.inv(ko.kriging_matrix(xyz, MODEL, stored))
    t_setup = time.perf_counter() - t1
    per_pt = max(1e-9, t_all - t_setup) / G.shape[0]
    return {"value": 1.0 / per_pt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle inverse x RHS on 10000 grid points (first 10 rows); steady per-point rate, "
                      "set-up (matrix + inv) %.2f s excluded; incl. set-up: %.0f points/s on this sample"
                      % (t_setup, G.shape[0] / t_all)}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import pykrige_b200 as pk
    from pykrige_b200 import multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL prints its version) must not write to stdout: rank 0 prints exactly ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: backend='cuda' has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    xyz, val, gx, gy = workload()
    gy_w = np.linspace(0.0, 1000.0 * world, GRID * world) if world > 1 else gy   # weak scaling: ny grows
    model = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model=MODEL, variogram_parameters=PARAMS)
    h = model._cuda_handle()
    stream = torch.cuda.Stream(device=dev)
    h.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    peak = None
    if rank == 0:
        peak = measure_fp64_peak(torch)

    d_gx = torch.from_numpy(gx).to(dev)
    d_gy = torch.from_numpy(gy).to(dev)
    d_gyw = torch.from_numpy(gy_w).to(dev)
    npt_weak = GRID * GRID * world
    first_w, count_w = multigpu.shard_range(npt_weak, rank, world)
    first_s, count_s = multigpu.shard_range(GRID * GRID, rank, world)
    d_out = torch.empty(2 * count_w, dtype=torch.float64, device=dev)

    def factor():
        # full re-assembly + re-factorisation every step (nothing cached), then the one broadcast
        model._kb_key = None
        return multigpu.prepare_sharded(model, dist if world > 1 else None)

    def step_dev(weak=True):
        with torch.cuda.stream(stream):
            flush.zero_()
        factor()
        if weak:
            h.execute_grid_dev(GRID, GRID * world, 1, d_gx.data_ptr(), d_gyw.data_ptr(), 0, 0, first_w, count_w,
                               d_out.data_ptr(), d_out.data_ptr() + 8 * count_w)
        else:
            h.execute_grid_dev(GRID, GRID, 1, d_gx.data_ptr(), d_gy.data_ptr(), 0, 0, first_s, count_s,
                               d_out.data_ptr(), d_out.data_ptr() + 8 * count_s)

    def step_e2e():
        with torch.cuda.stream(stream):
            flush.zero_()
        model._kb_key = None
        if world == 1:
            return model.execute("grid", gx, gy, backend="cuda")
        return multigpu.execute_grid_sharded(model, [gx, gy_w], dist)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)     # max over ranks
            dist.barrier()
        return float(ms.item()) / steps

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    h.reset_counters()
    ms_dev = timed(step_dev, args.steps)
    tm = h.timings()
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    launches = torch.tensor([tm["launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(launches)

    # end to end through the public API (host buffers)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # informational: the same grid through dtype='float64x' (fp64-class int8-slice tensor-core path)
    f64x = None
    if world == 1:
        try:
            model._kb_key = None
            zx, sx = model.execute("grid", gx, gy, backend="cuda", dtype="float64x")
            h.reset_counters()
            t0 = time.perf_counter()
            model._kb_key = None
            zx, sx = model.execute("grid", gx, gy, backend="cuda", dtype="float64x")
            wall = time.perf_counter() - t0
            tx = h.timings()
            model._kb_key = None
            z64, s64 = model.execute("grid", gx, gy, backend="cuda")
            f64x = {"e2e_points_per_s": GRID * GRID / wall, "solve_only_points_per_s": GRID * GRID / (tx["solve_ms"] * 1e-3),
                    "max_rel_dz_vs_float64": float(np.max(np.abs(zx - z64)) / np.max(np.abs(z64))),
                    "max_rel_dss_vs_float64": float(np.max(np.abs(sx - s64)) / np.max(np.abs(s64))),
                    "kernel": "solve_kernel_i8: tcgen05.mma kind::i8, 6x7-bit error-free slices, exact int32 accumulation in TMEM"}
        except Exception as e:  # noqa: BLE001
            f64x = {"error": "%s: %s" % (type(e).__name__, e)}

    strong = None
    if world > 1:
        step_dev(False)
        ms_strong = timed(lambda: step_dev(False), args.steps)
        strong = {"value": GRID * GRID / (ms_strong * 1e-3), "unit": UNIT, "ms_per_step": ms_strong,
                  "note": "fixed 1000x1000 grid split across the ranks (strong scaling)"}

    if rank == 0:
        value = npt_weak / (ms_dev * 1e-3)
        n1 = N_DATA + 1
        flop_pt = 2.0 * n1 * n1                                   # 2 n'^2 per grid point
        solve_ms = tm["solve_ms"] / args.steps                    # solve kernel launches of one step, this rank
        n_launch = max(1.0, tm["solve_launches"] / args.steps)
        achieved = count_w * flop_pt / (solve_ms * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "solve_kernel_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                traffic = None
        roofline = {
            "bound": "tensor", "kernel": "solve_kernel_pt (persistent point-tile kernel, fp64 DMMA mma.sync.m8n8k4; includes RHS generation and finalize)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "measured in this run: torch.matmul fp64 8192^3 (cuBLAS DGEMM), best of 3 "
                           "(MEASURED_PEAKS.json has no fp64 entry)",
            "algorithmic_flop_per_point": flop_pt, "points_per_launch": count_w / n_launch,
            "avg_launch_ms": solve_ms / n_launch,
            "note": "achieved uses the reference's algorithmic 2(N+1)^2 flop/point (inverse GEMV). The kernel "
                    "executes the covariance-form triangular product (~(N)^2 flop/point), so frac can exceed 1",
            "traffic": traffic,
        }
        cpu = cpu_baseline_sample() if world == 1 else None
        h2d = (3 * N_DATA + GRID + (GRID * world)) * 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world, {
                "grid_this_run": [GRID, GRID * world],
                "phases_ms_per_step_rank0": {k: tm[k] / args.steps for k in
                                             ("assemble_ms", "cholesky_ms", "trtri_ms", "pack_dual_ms", "solve_ms",
                                              "h2d_ms")},
                "solve_only_points_per_s_rank0": count_w / (solve_ms * 1e-3),
                "strong": strong,
                "float64x_int8_slices": f64x,
            }),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": npt_weak / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16 * npt_weak,
                    "api": "OrdinaryKriging.execute('grid', gx, gy, backend='cuda') -> kb200_set_problem + "
                           "kb200_execute_grid (host buffers), factorisation not cached"},
            "gpu_launches": int(launches.item()),
            "clocks": sampler.result(),
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
