import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np
from pykrige_b200 import _cabi
import pykrige_b200 as pk, cases
h = _cabi.aux_handle()
rng = np.random.default_rng(1)
X = rng.uniform(0, 1000, (60000, 2)); y = rng.normal(0, 1, 60000)
print(h.experimental_variogram(X, y, 6)[0].sum())
xyz, val = cases.synth_data(4242, 2000, 2)
for q in range(10): xyz[1999 - q] = xyz[2 * q]
m = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 250.0, 0.0], pseudo_inv=True)
print(m.execute("points", [10.0], [20.0], backend="cuda"))
