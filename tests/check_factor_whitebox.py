"""White-box check of the blocked Cholesky / triangular inverse against numpy through the debug taps of the C ABI (a GPU
tool, not collected by pytest: `python tests/check_factor_whitebox.py`; lives under tests/ because it uses the oracle)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pykrige_b200 as pk
from oracle import krige_oracle as ko
import os
os.environ["KB200_DEBUG"] = "1"
for n in (1000, 1500, 2500, 5000):
    xyz, val = cases.synth_data(3, n, 2)
    m = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
    h = m._ensure_problem("float64")
    t = h.timings()
    npad = (n + 255) // 256 * 256
    L = h.debug_fetch(1, npad * npad).reshape(npad, npad)
    W = h.debug_fetch(2, npad * npad).reshape(npad, npad)
    stored = ko.stored_parameters("exponential", [1.0, 300.0, 0.05])
    from scipy.spatial.distance import cdist
    d = cdist(xyz, xyz)
    C = 1.0 - ko.variogram("exponential", stored, d)
    np.fill_diagonal(C, 1.0)
    Lr = np.linalg.cholesky(C)
    Ld = np.tril(L[:n, :n])
    err = np.abs(Ld - Lr)
    bc = [float(err[:, j:j + 64].max()) for j in range(0, n, 64)]
    print(n, "launches", t["launches"], "chol_ms", t["cholesky_ms"], "max err per block column", ["%.1e" % e for e in bc], flush=True)
    Wd = np.tril(W[:n, :n])
    print("   W err", float(np.abs(Wd - np.linalg.inv(Lr)).max()), flush=True)
