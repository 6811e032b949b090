"""Driver for the round-2 ncu captures: one cfg2-sized problem (N=5000, exponential) kriged on a short point list
through each arithmetic (float64 DMMA, float32 tcgen05 TF32, float64x tcgen05 INT8 slices); the factorisation
kernels run once per dtype. Usage: python scripts/ncu_r02_drive.py [dtype ...] [--m M]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pykrige_b200 as pk

args = [a for a in sys.argv[1:] if not a.startswith("--")]
m = 148 * 128 * 2
for a in sys.argv[1:]:
    if a.startswith("--m="):
        m = int(a[4:])
n = 5000
for a in sys.argv[1:]:
    if a.startswith("--n="):
        n = int(a[4:])
dtypes = args or ["float64", "float32", "float64x"]
xyz, val = cases.synth_data(1002, n, 2)
mdl = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
rng = np.random.default_rng(1)
px = rng.uniform(0, 1000, m); py = rng.uniform(0, 1000, m)
for dt in dtypes:
    mdl._kb_key = None
    z, ss = mdl.execute("points", px, py, backend="cuda", dtype=dt)
    t = mdl._kb_handle.timings()
    print(dt, m, "solve_ms", t["solve_ms"], float(z[0]), float(ss[0]), flush=True)
