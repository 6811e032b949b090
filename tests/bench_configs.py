"""Time the other BASELINE.json configs on ONE GPU (full data size, points sharded as one slice) and
spot-check parity against the CPU oracle on a subsample. Prints one JSON line per config.
    python tests/bench_configs.py [cfg1 cfg3 cfg4 cfg5 ...] [--frac F]   (F = fraction of the grid to krige)
"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # lives under tests/: it checks parity against the oracle
import cases  # noqa: E402
import pykrige_b200 as pk  # noqa: E402
from oracle import krige_oracle as ko  # noqa: E402

R = 1e-5


def parity(z, zo, s, so):
    ez = float(np.max(np.abs(z - zo)) / np.max(np.abs(zo)))
    es = float(np.max(np.abs(s - so)) / np.max(np.abs(so)))
    ok = bool(np.allclose(z, zo, rtol=R, atol=R * np.abs(zo).max()) and np.allclose(s, so, rtol=R, atol=R * np.abs(so).max()))
    return {"max_rel_z": ez, "max_rel_ss": es, "pass_rtol_1e-5": ok}


def run(name, frac):
    t_all = time.perf_counter()
    if name == "cfg1":
        xyz, val = cases.synth_data(1001, 100, 2)
        axes = [np.linspace(0, 1000, 50), np.linspace(0, 1000, 50)]
        model = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="spherical", variogram_parameters=[1.0, 400.0, 0.05])
        okw = dict(model="spherical", stored=[0.95, 400.0, 0.05]); k = None; flop_pt = 2.0 * 101**2
    elif name == "cfg3":
        xyz, val = cases.synth_data(1003, 8000, 3)
        axes = [np.linspace(0, 1000, 200), np.linspace(0, 1000, 200), np.linspace(0, 250, 50)]
        model = pk.OrdinaryKriging3D(xyz[:, 0], xyz[:, 1], xyz[:, 2], val, variogram_model="gaussian", variogram_parameters=[1.0, 300.0, 0.05])
        okw = dict(model="gaussian", stored=[0.95, 300.0, 0.05]); k = None; flop_pt = 2.0 * 8001**2
    elif name == "cfg4":
        xyz, val = cases.synth_data(1004, 10000, 2)
        axes = [np.linspace(0, 1000, 2000), np.linspace(0, 1000, 2000)]
        model = pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05],
                                    drift_terms=["regional_linear"])
        okw = dict(model="exponential", stored=[0.95, 300.0, 0.05], regional_linear=True); k = None; flop_pt = 2.0 * 10003**2
    elif name == "cfg5":
        xyz, val = cases.synth_data(1005, 100000, 2)
        axes = [np.linspace(0, 1000, 4000), np.linspace(0, 1000, 4000)]
        model = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 50.0, 0.05])
        okw = dict(model="exponential", stored=[0.95, 50.0, 0.05]); k = 64; flop_pt = (2.0 / 3.0) * 65**3 + 2.0 * 65**2
    else:
        raise SystemExit("unknown config " + name)
    dim = xyz.shape[1]
    npt = int(np.prod([a.size for a in axes]))
    count = max(1, int(npt * frac))
    h = model._ensure_problem("float64", knn=k is not None)     # factor (timed separately below)
    h.reset_counters()
    t0 = time.perf_counter()
    model._kb_key = None
    h = model._ensure_problem("float64", knn=k is not None)
    t_factor = time.perf_counter() - t0
    gz = axes[2] if dim == 3 else None
    # warm + timed slice of the flattened grid through the host-buffer C ABI
    if k is None:
        h.execute_grid(axes[0], axes[1], gz, None, 0, min(count, 65536))
        h.reset_counters()
        t0 = time.perf_counter()
        z, ss = h.execute_grid(axes[0], axes[1], gz, None, 0, count)
    else:
        h.execute_knn_grid(k, axes[0], axes[1], gz, 0, min(count, 65536))
        h.reset_counters()
        t0 = time.perf_counter()
        z, ss = h.execute_knn_grid(k, axes[0], axes[1], gz, 0, count)
    t_exec = time.perf_counter() - t0
    tm = h.timings()
    # oracle on a subsample of the slice + exact hits
    rng = np.random.default_rng(7)
    m_chk = 256 if xyz.shape[0] > 20000 else 1024
    pick = np.sort(rng.choice(count, size=min(m_chk, count), replace=False))
    G = ko.grid_points(axes)[pick] if npt <= 5_000_000 else None
    if G is None:
        nx, ny = axes[0].size, axes[1].size
        ix, iy = pick % nx, (pick // nx) % ny
        cols = [axes[0][ix], axes[1][iy]] + ([axes[2][pick // (nx * ny)]] if dim == 3 else [])
        G = np.column_stack(cols)
    if k is None:
        zo, so = ko.krige_chunked(xyz, val, okw["model"], okw["stored"], G, regional_linear=okw.get("regional_linear", False))
    else:
        zo, so = ko.krige(xyz, val, okw["model"], okw["stored"], G, n_closest_points=k)
    par = parity(z[pick], zo, ss[pick], so)
    dev_ms = tm["solve_ms"] if k is None else tm["knn_solve_ms"]
    out = {"config": name, "n": int(xyz.shape[0]), "grid_points_total": npt, "points_timed": count,
           "factor_wall_s": t_factor, "execute_wall_s": t_exec,
           "points_per_s_e2e_host_buffers": count / t_exec, "points_per_s_device_kernel": count / (dev_ms * 1e-3),
           "algorithmic_tflops": count * flop_pt / (dev_ms * 1e-3) / 1e12,
           "timings_ms": {kk: tm[kk] for kk in tm if tm[kk]}, "parity_vs_oracle": par,
           "wall_total_s": time.perf_counter() - t_all}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    frac = 1.0
    for i, a in enumerate(sys.argv):
        if a == "--frac":
            frac = float(sys.argv[i + 1])
            args = [x for x in args if x != sys.argv[i + 1]]
    for name in (args or ["cfg1", "cfg3", "cfg4", "cfg5"]):
        run(name, frac)
