"""One moving-window launch (cfg5 data: N=100000, k=64) on a short slice of the 4000x4000 grid, for ncu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pykrige_b200 as pk
m = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8 * 16
xyz, val = cases.synth_data(1005, 100000, 2)
ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 50.0, 0.05])
g = np.linspace(0, 1000, 4000)
h = ok._ensure_problem("float64", knn=True)
z, ss = h.execute_knn_grid(64, g, g, None, 4000 * 1000, m)
h.reset_counters()
z, ss = h.execute_knn_grid(64, g, g, None, 4000 * 2000, m)
t = h.timings()
print("knn", m, "points", t["knn_solve_ms"], "ms", m / (t["knn_solve_ms"] * 1e-3), "points/s", float(z[0]), float(ss[0]))
