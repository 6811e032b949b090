import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cases, pykrige_b200 as pk
for n in (5000, 10000):
    xyz, val = cases.synth_data(1, n, 2)
    ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
    ok.execute("points", [1.0], [2.0], backend="cuda")
    h = ok._kb_handle
    for rep in range(2):
        h.reset_counters(); ok._kb_key = None
        ok.execute("points", np.linspace(0, 1000, 64), np.linspace(0, 1000, 64), backend="cuda")
    t = h.timings(); print(n, {k: round(v, 3) for k, v in t.items() if v}, flush=True)
