"""fp64 solve kernel on 125 000 points of config 2 (what one of 8 GPUs gets): 64- vs 48-point tiles are chosen by the library;
prints the solve time of a 125 000-point and of a 132 608-point (= 14 x 148 x 64) slice."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pykrige_b200 as pk
xyz, val = cases.synth_data(1002, 5000, 2)
ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
g = np.linspace(0, 1000, 1000)
h = ok._ensure_problem()
for count in (125000, 132608, 125000, 1000000):
    h.execute_grid(g, g, None, None, 0, count)
    h.reset_counters()
    h.execute_grid(g, g, None, None, 0, count)
    t = h.timings()
    print(count, "points", round(t["solve_ms"], 2), "ms", round(count / (t["solve_ms"] * 1e-3)), "points/s", flush=True)
