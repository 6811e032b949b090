"""One general-path (indefinite) factorisation for an ncu launch list / compute-sanitizer: hole-effect, N = argv[1] (1900)."""
import os, sys, numpy as np
os.environ["KB200_GJ"] = "blocked"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cases, pykrige_b200 as pk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
xyz, val = cases.synth_data(21, n, 2)
ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="hole-effect", variogram_parameters=[1.0, 250.0, 0.02])
z, ss = ok.execute("points", np.linspace(0, 1000, 64), np.linspace(0, 1000, 64), backend="cuda")
print(float(z[0]), float(ss[0]))
