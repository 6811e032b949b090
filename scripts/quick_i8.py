"""fp64-class int8-slice kernel (dtype='float64x') vs the fp64 DMMA kernel: agreement and rate."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pykrige_b200 as pk

for n, m, cls in ((300, 1000, "ok"), (1000, 20000, "uk"), (5000, 400000, "ok")):
    xyz, val = cases.synth_data(1002, n, 2)
    if cls == "ok":
        mdl = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
    else:
        mdl = pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05],
                                  drift_terms=["regional_linear"])
    rng = np.random.default_rng(1)
    px = np.concatenate([rng.uniform(0, 1000, m), xyz[:8, 0]]); py = np.concatenate([rng.uniform(0, 1000, m), xyz[:8, 1]])
    out = {}
    for dt in ("float64", "float64x"):
        mdl._kb_key = None
        z, ss = mdl.execute("points", px, py, backend="cuda", dtype=dt)
        h = mdl._kb_handle
        h.reset_counters()
        z, ss = mdl.execute("points", px, py, backend="cuda", dtype=dt)
        out[dt] = (z, ss, h.timings()["solve_ms"])
    z64, s64 = out["float64"][:2]; zx, sx = out["float64x"][:2]
    print(json.dumps({"n": n, "m": int(px.size), "f64_pts_per_s": px.size / (out["float64"][2] * 1e-3),
                      "f64x_pts_per_s": px.size / (out["float64x"][2] * 1e-3),
                      "max_rel_z": float(np.abs(zx - z64).max() / np.abs(z64).max()),
                      "max_rel_ss": float(np.abs(sx - s64).max() / np.abs(s64).max()),
                      "ss_at_hits": [float(v) for v in sx[-3:]]}), flush=True)
