"""CPU emulation of the C ABI (include/krige_b200.h) — TEST INFRASTRUCTURE ONLY.

The host wrappers (pykrige_b200/ok.py, uk.py, ok3d.py, uk3d.py, _base.py, multigpu.py) talk to the device through
`_cabi.Handle`. A box without a GPU cannot create one (kb200_create returns KB200_ECUDA: no CPU fallback), so the
host logic between the reference-facing API and the C ABI — point planning, mask compaction, drift columns and their
order, what is handed over in ORIGINAL vs ADJUSTED coordinates, device-drift configuration, output scatter and
shaping — would only ever run on the GPU box. `EmulatedHandle` implements the DOCUMENTED semantics of every entry
point the wrappers call, argument for argument as the header states them, with the CPU oracle's reference
formulation (oracle/krige_oracle.py: gamma-form matrix, explicit inverse, inverse x RHS); tests monkeypatch it in
place of `_cabi.Handle` and compare whole `execute()` calls with the fixtures of the imported reference.
It is never importable from the product package, and nothing here is timed or shipped."""
import numpy as np

from oracle import krige_oracle as ko

MODEL_NAMES = {0: "linear", 1: "power", 2: "gaussian", 3: "exponential", 4: "spherical", 5: "hole-effect"}
TABLE_MODEL = 6


class EmulatedHandle:
    """Same method surface as pykrige_b200._cabi.Handle (the subset the class wrappers and multigpu use)."""

    def __init__(self, device=-1):
        self.geo = False
        self.pinv = False
        self.table = None
        self.wells = None
        self.ext = None
        self.problem = None
        self.calls = []

    def close(self):
        pass

    # ---- configuration that precedes a problem description --------------------------------------------------
    def set_coordinates(self, geographic):
        self.geo = bool(geographic)
        self.problem = None

    def set_pseudo_inverse(self, enable):
        self.pinv = bool(enable)
        self.problem = None

    def set_variogram_table(self, nodes, dmax):
        nodes = np.asarray(nodes, dtype=np.float64)
        if nodes.size < 16 or not np.all(np.isfinite(nodes)) or not dmax > 0:
            raise ValueError("variogram table: 16 <= n_nodes, finite nodes, dmax > 0")
        self.table = (nodes.copy(), float(dmax))
        self.problem = None

    def set_device_drift(self, wells, ext):
        self.wells = None if wells is None or len(wells) == 0 else np.asarray(wells, dtype=np.float64).reshape(-1, 3).copy()
        self.ext = None if ext is None else tuple(np.asarray(a, dtype=np.float64).copy() for a in ext)

    def set_stream(self, stream):
        pass

    def timings(self):
        return {"launches": 1.0, "solve_launches": 1.0}

    def reset_counters(self):
        pass

    # ---- the variogram the following problem uses -----------------------------------------------------------
    def _model(self, model, vparams):
        if model == TABLE_MODEL:
            if self.table is None:
                raise RuntimeError("KB200_VG_TABLE: call kb200_set_variogram_table first")
            nodes, dmax = self.table
            t_nodes = np.linspace(0.0, 1.0, nodes.size)          # node i sits at sqrt(d / dmax) = i / (n - 1)

            def gamma(m, d):
                d = np.asarray(d, dtype=np.float64)
                if np.any(d > dmax * (1 + 1e-12)):
                    raise AssertionError("distance %g beyond the tabulated range %g" % (float(np.max(d)), dmax))
                return np.interp(np.sqrt(np.clip(d, 0.0, dmax) / dmax), t_nodes, nodes)
            return gamma, []
        if model not in MODEL_NAMES:
            raise NotImplementedError("variogram model has no device implementation")
        want = 2 if model == 0 else 3
        if len(vparams) != want:
            raise ValueError("wrong number of variogram parameters")
        return MODEL_NAMES[model], [float(v) for v in vparams]

    def _adjust(self, P):
        """adjusted = Mt @ (p - center) + center (krige_b200.h: kb200_set_problem)."""
        p = self.problem
        if p["geo"]:
            return P
        return (P - p["center"]) @ p["Mt"].T + p["center"]

    def _describe(self, knn, dim, x, y, z, values, center, aniso, model, vparams, exact_values, eps, n_rl, drift_data):
        if dim not in (2, 3):
            raise ValueError("dim must be 2 or 3")
        cols = [np.asarray(x, float), np.asarray(y, float)] + ([np.asarray(z, float)] if dim == 3 else [])
        n = cols[0].size
        hd = [] if drift_data is None else [np.asarray(c, float).ravel() for c in drift_data]
        if self.geo and (dim != 2 or n_rl or hd):
            raise NotImplementedError("geographic coordinates: ordinary kriging in two dimensions only")
        if n_rl not in (0, dim):
            raise ValueError("n_rl must be 0 or dim")
        if knn and (n_rl or hd):
            raise NotImplementedError("moving window supports ordinary kriging only")
        n_dev = (0 if self.wells is None else self.wells.shape[0]) + (0 if self.ext is None else 1)
        if n_dev > len(hd):
            raise ValueError("device drift terms exceed the described drift columns")
        if n_dev and dim != 2:
            raise NotImplementedError("point_log / external_Z drift terms are two-dimensional")
        for c in hd:
            assert c.size == n, "drift_data is column-major n x n_hd"
        fn, m = self._model(int(model), list(np.ravel(vparams)) if vparams is not None else [])
        self.problem = dict(knn=knn, dim=dim, geo=self.geo, pinv=self.pinv and not knn, X=np.column_stack(cols),
                            values=np.asarray(values, float).copy(), center=np.asarray(center, float)[:dim].copy(),
                            Mt=np.asarray(aniso, float).reshape(dim, dim).copy(), fn=fn, m=m,
                            exact=bool(exact_values), eps=float(eps), n_rl=int(n_rl), hd=hd, n_dev=n_dev,
                            wells=self.wells, ext=self.ext, a=None)
        assert abs(float(eps) - ko.EPS) < 1e-30, "the oracle hard-codes the reference's eps"

    def set_problem(self, dim, dtype, x, y, z, values, center, aniso, model, vparams, exact_values, eps,
                    n_rl=0, drift_data=None):
        self.calls.append("set_problem")
        self._describe(False, dim, x, y, z, values, center, aniso, model, vparams, exact_values, eps, n_rl, drift_data)
        p = self.problem
        if p["pinv"] and dtype != 0:
            raise NotImplementedError("pseudo_inv=True runs in float64 only")
        if p["geo"]:
            self.ready = True
            try:
                import torch
                self.blob_t = torch.ones(1, dtype=torch.float64)
            except ImportError:
                self.blob_t = None
            return
        P = self._adjust(p["X"])
        dcols = [P[:, c] for c in range(dim)] if p["n_rl"] else []
        p["P"] = P
        p["a"] = ko.kriging_matrix(P, p["fn"], p["m"], dcols + p["hd"])
        if not p["pinv"]:
            # the factorisation reports a singular system here (scipy.linalg.inv raises in the reference's execute)
            import scipy.linalg
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                scipy.linalg.inv(p["a"])
        try:
            import torch
            self.blob_t = torch.from_numpy(p["a"].ravel().copy())
        except ImportError:
            self.blob_t = None
        self.ready = True

    # ---- multi-GPU factor blob: here simply the dense kriging matrix (kb200_describe_problem / kb200_blob_commit) ----
    def describe_problem(self, dim, dtype, x, y, z, values, center, aniso, model, vparams, exact_values, eps,
                         n_rl=0, drift_data=None):
        """No 'device work': records the description and allocates the blob a broadcast will fill."""
        import torch
        self.calls.append("describe_problem")
        self._describe(False, dim, x, y, z, values, center, aniso, model, vparams, exact_values, eps, n_rl, drift_data)
        p = self.problem
        nt = p["X"].shape[0] + (dim if p["n_rl"] else 0) + len(p["hd"]) + 1
        self.blob_t = torch.zeros(1 if p["geo"] else nt * nt, dtype=torch.float64)
        self.ready = False

    def blob_commit(self):
        self.calls.append("blob_commit")
        p = self.problem
        if not p["geo"]:
            p["P"] = self._adjust(p["X"])
            nt = int(round(self.blob_t.numel() ** 0.5))
            p["a"] = self.blob_t.numpy().reshape(nt, nt).copy()
            assert np.any(p["a"] != 0.0), "blob_commit before the broadcast arrived"
        self.ready = True

    def set_problem_knn(self, dim, x, y, z, values, center, aniso, model, vparams, exact_values, eps):
        self.calls.append("set_problem_knn")
        self._describe(True, dim, x, y, z, values, center, aniso, model, vparams, exact_values, eps, 0, None)
        if not self.problem["geo"]:
            self.problem["P"] = self._adjust(self.problem["X"])

    # ---- device-evaluated drift terms at the prediction points (kb200_set_device_drift) ----------------------
    def _device_drift_columns(self, Q_orig, Q_adj):
        p = self.problem
        cols = []
        if p["wells"] is not None:
            for wx, wy, strength in p["wells"]:
                with np.errstate(divide="ignore"):
                    ld = np.log(np.sqrt((Q_adj[:, 0] - wx) ** 2 + (Q_adj[:, 1] - wy) ** 2))
                ld[np.isinf(ld)] = -100.0                                      # uk.py:960-963
                cols.append(-strength * ld)
        if p["ext"] is not None:
            ax, ay, Zg = p["ext"]
            cols.append(_bilinear(ax, ay, Zg, Q_orig[:, 0], Q_orig[:, 1]))      # sampled at ORIGINAL coordinates
        return cols

    def _krige(self, Q_orig, drift_pts):
        p = self.problem
        assert p is not None and not p["knn"], "describe the global problem first"
        assert getattr(self, "ready", False), "execute before kb200_set_problem / kb200_blob_commit"
        Q_orig = np.asarray(Q_orig, dtype=np.float64)
        m = Q_orig.shape[0]
        if m == 0:
            return np.zeros(0), np.zeros(0)
        if p["geo"]:
            assert drift_pts is None
            return ko.krige_geographic(p["X"], p["values"], p["fn"], p["m"], Q_orig, exact_values=p["exact"])
        Q = self._adjust(Q_orig)
        pcols = [Q[:, c] for c in range(p["dim"])] if p["n_rl"] else []
        pcols += self._device_drift_columns(Q_orig, Q)
        n_host = len(p["hd"]) - p["n_dev"]
        if n_host:
            d = np.asarray(drift_pts, dtype=np.float64)
            assert d.shape == (n_host, m), "drift_pts is column-major m x (n_hd - n_dev): got %r" % (d.shape,)
            pcols += [d[c] for c in range(n_host)]
        else:
            assert drift_pts is None or np.size(drift_pts) == 0
        pinv = "pinv" if p["pinv"] else None     # pinv and pinvh agree on a symmetric matrix; one ABI switch covers both
        return ko.exec_vector(p["a"], p["P"], Q, p["values"], p["fn"], p["m"], p["exact"], pcols, pinv)

    def _grid(self, gx, gy, gz):
        axes = [np.asarray(gx, float), np.asarray(gy, float)] + ([np.asarray(gz, float)] if gz is not None else [])
        assert len(axes) == self.problem["dim"], "gz NULL and nz = 1 for 2-D"
        return ko.grid_points(axes)

    # ---- execute ------------------------------------------------------------------------------------------
    def execute_points(self, px, py, pz=None, drift_pts=None):
        self.calls.append("execute_points")
        cols = [np.asarray(px, float), np.asarray(py, float)] + ([np.asarray(pz, float)] if pz is not None else [])
        assert len(cols) == self.problem["dim"]
        return self._krige(np.column_stack(cols), drift_pts)

    def execute_grid(self, gx, gy, gz=None, drift_pts=None, first=0, count=None):
        self.calls.append("execute_grid")
        G = self._grid(gx, gy, gz)
        count = G.shape[0] - first if count is None else count
        assert 0 <= first and first + count <= G.shape[0]
        return self._krige(G[first:first + count], drift_pts)

    def _knn(self, k, Q_orig):
        p = self.problem
        assert p is not None and p["knn"], "kb200_set_problem_knn first"
        if Q_orig.shape[0] == 0:
            return np.zeros(0), np.zeros(0)
        if k > p["X"].shape[0]:
            raise ValueError("n_closest_points exceeds the number of data points")
        try:
            if p["geo"]:
                return ko.krige_geographic(p["X"], p["values"], p["fn"], p["m"], Q_orig, exact_values=p["exact"],
                                           n_closest_points=int(k))
            return ko.exec_moving_window(p["P"], self._adjust(Q_orig), p["values"], p["fn"], p["m"], int(k), p["exact"])
        except np.linalg.LinAlgError:
            raise ValueError("Singular matrix")                               # cok.pyx:176-179

    def execute_knn_points(self, k, px, py, pz=None):
        self.calls.append("execute_knn_points")
        cols = [np.asarray(px, float), np.asarray(py, float)] + ([np.asarray(pz, float)] if pz is not None else [])
        return self._knn(k, np.column_stack(cols))

    def execute_knn_grid(self, k, gx, gy, gz=None, first=0, count=None):
        self.calls.append("execute_knn_grid")
        G = self._grid(gx, gy, gz)
        count = G.shape[0] - first if count is None else count
        return self._knn(k, G[first:first + count])

    def statistics(self, n):
        raise NotImplementedError("the emulator has no factor to read the statistics from (host loop instead)")


def _bilinear(ax, ay, Zg, x, y):
    """The raster sampler of uk.py:512-628 restated: along each axis the bracketing nodes are i1 = the LAST index whose
    coordinate is <= the query and i2 = the FIRST index whose coordinate is >= it (for ascending axes the enclosing cell;
    for descending or unsorted axes whatever that rule selects — the reference does not sort, and neither does the
    device); on a node the node value, on a grid line the linear interpolation along the other axis, else the bilinear
    form over the two bracketing nodes per axis."""
    ax, ay, x, y = (np.asarray(a, dtype=np.float64) for a in (ax, ay, x, y))

    def bracket(a, v):
        le = a[None, :] <= v[:, None]
        ge = a[None, :] >= v[:, None]
        assert np.all(le.any(axis=1)) and np.all(ge.any(axis=1)), "the raster does not cover the query"
        i1 = a.size - 1 - np.argmax(le[:, ::-1], axis=1)
        i2 = np.argmax(ge, axis=1)
        return i1, i2
    x1, x2 = bracket(ax, x)
    y1, y2 = bracket(ay, y)
    wx1, wx2 = ax[x2] - x, x - ax[x1]
    wy1, wy2 = ay[y2] - y, y - ay[y1]
    dx, dy = ax[x2] - ax[x1], ay[y2] - ay[y1]
    with np.errstate(divide="ignore", invalid="ignore"):
        full = (Zg[y1, x1] * wx1 * wy1 + Zg[y1, x2] * wx2 * wy1 + Zg[y2, x1] * wx1 * wy2 + Zg[y2, x2] * wx2 * wy2) / (dx * dy)
        on_row = (Zg[y1, x1] * wx1 + Zg[y1, x2] * wx2) / dx          # y sits on a grid line
        on_col = (Zg[y1, x1] * wy1 + Zg[y2, x1] * wy2) / dy          # x sits on a grid line
    return np.where(y1 == y2, np.where(x1 == x2, Zg[y1, x1], on_row), np.where(x1 == x2, on_col, full))
