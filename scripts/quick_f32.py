"""Quick fp32 (tcgen05 TF32) sanity + timing vs the fp64 path on one GPU."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pykrige_b200 as pk

for n, m in ((1000, 20000), (10000, 400000)):
    xyz, val = cases.synth_data(1004, n, 2)
    uk = pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential",
                             variogram_parameters=[1.0, 300.0, 0.05], drift_terms=["regional_linear"])
    rng = np.random.default_rng(1)
    px, py = rng.uniform(0, 1000, m), rng.uniform(0, 1000, m)
    out = {}
    for dt in ("float64", "float32"):
        uk._kb_key = None
        z, ss = uk.execute("points", px, py, backend="cuda", dtype=dt)     # includes factorisation
        h = uk._kb_handle
        h.reset_counters()
        t0 = time.perf_counter()
        z, ss = uk.execute("points", px, py, backend="cuda", dtype=dt)
        dtm = time.perf_counter() - t0
        out[dt] = (z, ss, dtm, h.timings()["solve_ms"])
    z64, s64 = out["float64"][:2]; z32, s32 = out["float32"][:2]
    print(json.dumps({"n": n, "m": m, "f64_pts_per_s": m / (out["float64"][3] * 1e-3), "f32_pts_per_s": m / (out["float32"][3] * 1e-3),
                      "max_rel_z": float(np.abs(z32 - z64).max() / np.abs(z64).max()),
                      "max_rel_ss": float(np.abs(s32 - s64).max() / np.abs(s64).max()),
                      "f32_algorithmic_tflops": m * 2.0 * (n + 3) ** 2 / (out["float32"][3] * 1e-3) / 1e12}), flush=True)
