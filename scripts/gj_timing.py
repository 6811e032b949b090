"""General (indefinite) path: blocked Gauss-Jordan vs the column-at-a-time form, hole-effect on dense 2-D scatter."""
import os, sys, json, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cases, pykrige_b200 as pk
for n in (1900, 5000):
    xyz, val = cases.synth_data(21, n, 2)
    pts = cases.synth_points(21, 2000, 2, xyz)
    res = {}
    for mode in ("blocked", "scalar"):
        os.environ["KB200_GJ"] = mode
        ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="hole-effect", variogram_parameters=[1.0, 250.0, 0.02])
        ok.execute("points", pts[:4, 0], pts[:4, 1], backend="cuda")
        h = ok._kb_handle
        h.reset_counters(); ok._kb_key = None
        z, ss = ok.execute("points", pts[:, 0], pts[:, 1], backend="cuda")
        t = h.timings()
        res[mode] = (z, ss)
        print(json.dumps({"n": n, "form": mode, "factor_ms": round(t["cholesky_ms"], 3), "launches": t["launches"]}), flush=True)
    dz = np.max(np.abs(res["blocked"][0] - res["scalar"][0])) / np.max(np.abs(res["scalar"][0]))
    ds = np.max(np.abs(res["blocked"][1] - res["scalar"][1])) / np.max(np.abs(res["scalar"][1]))
    print(json.dumps({"n": n, "blocked_vs_scalar_max_rel_z": dz, "max_rel_ss": ds}), flush=True)
