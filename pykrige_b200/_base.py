"""Shared host logic of the four kriging classes (not part of the reference's API surface).

The reference repeats this logic in ok.py / uk.py / ok3d.py / uk3d.py; here it lives once:
variogram model selection (ok.py:208-253), the ``backend='cuda'`` dispatch that replaces the
``backend`` string switch of ``execute`` (ok.py:971-1010), point-list / grid / mask handling
(ok.py:842-900, ok3d.py:833-898) and output shaping (ok.py:1012-1020).
"""
import warnings
import numpy as np

from . import variogram_models
from . import core
from . import _cabi


class KrigeBase:
    eps = 1.0e-10  # cutoff for comparison to zero (ok.py:177)
    variogram_dict = {
        "linear": variogram_models.linear_variogram_model,
        "power": variogram_models.power_variogram_model,
        "gaussian": variogram_models.gaussian_variogram_model,
        "spherical": variogram_models.spherical_variogram_model,
        "exponential": variogram_models.exponential_variogram_model,
        "hole-effect": variogram_models.hole_effect_variogram_model,
    }
    _ndim = 2
    _backend_name = "ordinary kriging"

    # ---- variogram model selection (ok.py:208-253; GSTools models arrive as 'custom') ----
    def _select_variogram(self, variogram_model, variogram_function, gstools_dim_ok):
        self.variogram_model = variogram_model
        self.model = None
        overrides = {}
        if hasattr(self.variogram_model, "pykrige_kwargs"):
            self.model = self.variogram_model
            gstools_dim_ok(self.model)
            self.variogram_model = "custom"
            variogram_function = self.model.pykrige_vario
            overrides["variogram_parameters"] = []
            overrides["gstools"] = self.model
        if self.variogram_model not in self.variogram_dict.keys() and self.variogram_model != "custom":
            raise ValueError("Specified variogram model '%s' is not supported." % variogram_model)
        elif self.variogram_model == "custom":
            if variogram_function is None or not callable(variogram_function):
                raise ValueError("Must specify callable function for custom variogram model.")
            self.variogram_function = variogram_function
        else:
            self.variogram_function = self.variogram_dict[self.variogram_model]
        return overrides

    def _print_variogram(self):
        p = self.variogram_model_parameters
        if self.variogram_model == "linear":
            print("Using '%s' Variogram Model" % "linear")
            print("Slope:", p[0])
            print("Nugget:", p[1], "\n")
        elif self.variogram_model == "power":
            print("Using '%s' Variogram Model" % "power")
            print("Scale:", p[0])
            print("Exponent:", p[1])
            print("Nugget:", p[2], "\n")
        elif self.variogram_model == "custom":
            print("Using Custom Variogram Model")
        else:
            print("Using '%s' Variogram Model" % self.variogram_model)
            print("Partial Sill:", p[0])
            print("Full Sill:", p[0] + p[2])
            print("Range:", p[1])
            print("Nugget:", p[2], "\n")

    # ---- experimental variogram: computed on first access when the parameters were given explicitly
    #      (the reference always runs the O(N^2) pdist in the constructor, core.py:432-436) -------------
    def _get_lags(self):
        if callable(self._lags):
            self._lags, self._semivariance = self._lags()
        return self._lags

    def _set_lags(self, v):
        self._lags = v

    def _get_semivariance(self):
        self._get_lags()
        return self._semivariance

    def _set_semivariance(self, v):
        if v is not None or not callable(getattr(self, "_lags", None)):
            self._semivariance = v

    lags = property(_get_lags, _set_lags)
    semivariance = property(_get_semivariance, _set_semivariance)

    # ---- cross-validation statistics: lazy (the reference runs this O(N^4) loop in the
    #      constructor of OK3D/UK/UK3D, ok3d.py:352, uk.py:380, uk3d.py:380) -----
    def _stats_inputs(self):
        raise NotImplementedError

    def _device_statistics(self):
        """delta, sigma, epsilon from the Cholesky factor of the device problem (kb200_statistics,
        csrc/variogram.cu: O(N) after the factorisation instead of the reference's N solves). Returns
        None when this problem has no device twin (custom variogram, pseudo_inv, indefinite
        covariance form, no CUDA device) — the caller then runs the reference's loop on the host."""
        if not _cabi.device_available() or getattr(self, "pseudo_inv", False):
            return None                          # core._krige solves with lstsq under pseudo_inv (core.py:749-750)
        try:
            key = getattr(self, "_kb_key", None)
            if key is not None and key[1] is False and key == self._problem_signature(key[0], False):
                h = self._cuda_handle()           # the factor of the last global execute() is still there
            else:
                h = self._ensure_problem("float64")
            delta, sigma = h.statistics(len(self._stats_inputs()[1]))
        except (NotImplementedError, _cabi.KrigeB200Error, np.linalg.LinAlgError, ValueError, MemoryError):
            return None
        keep = (sigma * sigma >= core.eps) & (sigma > core.eps)      # core.py:818-819, 829-831
        delta, sigma = delta[keep], sigma[keep]
        return delta, sigma, delta / sigma

    def _compute_statistics(self, device="auto"):
        X, y = self._stats_inputs()
        res = self._device_statistics() if device in ("auto", True) else None
        if res is None:
            if device is True:
                raise _cabi.KrigeB200Error("cross-validation statistics: this problem has no device route")
            res = core._find_statistics(
                X, y, self.variogram_function, self.variogram_model_parameters,
                getattr(self, "coordinates_type", "euclidean"), getattr(self, "pseudo_inv", False),
            )
        self._delta, self._sigma, self._epsilon = res
        self._Q1 = core.calcQ1(self._epsilon)
        self._Q2 = core.calcQ2(self._epsilon)
        self._cR = core.calc_cR(self._Q2, self._sigma)
        self._stats_state = "done"

    def _stat(self, name):
        state = getattr(self, "_stats_state", "off")
        if state == "off":
            return None
        if state == "lazy":
            self._compute_statistics()
        return getattr(self, "_" + name)

    delta = property(lambda self: self._stat("delta"))
    sigma = property(lambda self: self._stat("sigma"))
    epsilon = property(lambda self: self._stat("epsilon"))
    Q1 = property(lambda self: self._stat("Q1"))
    Q2 = property(lambda self: self._stat("Q2"))
    cR = property(lambda self: self._stat("cR"))

    # ---- small public helpers kept from the reference API (ok.py:555-624) ---------------
    def display_variogram_model(self):
        """Displays variogram model with the actual binned data."""
        import matplotlib.pyplot as plt

        fig = plt.figure()
        ax = fig.add_subplot(111)
        ax.plot(self.lags, self.semivariance, "r*")
        ax.plot(self.lags, self.variogram_function(self.variogram_model_parameters, self.lags), "k-")
        plt.show()

    def get_variogram_points(self):
        """Returns both the lags and the variogram function evaluated at each of them."""
        return self.lags, self.variogram_function(self.variogram_model_parameters, self.lags)

    def switch_verbose(self):
        self.verbose = not self.verbose

    def switch_plotting(self):
        self.enable_plotting = not self.enable_plotting

    def get_epsilon_residuals(self):
        return self.epsilon

    def plot_epsilon_residuals(self):
        import matplotlib.pyplot as plt

        fig = plt.figure()
        ax = fig.add_subplot(111)
        ax.scatter(range(self.epsilon.size), self.epsilon, c="k", marker="*")
        ax.axhline(y=0.0)
        plt.show()

    def get_statistics(self):
        return self.Q1, self.Q2, self.cR

    def print_statistics(self):
        print("Q1 =", self.Q1)
        print("Q2 =", self.Q2)
        print("cR =", self.cR)

    # ---- the backend='cuda' arm ---------------------------------------------------------
    TABLE_MODEL_ID = 6          # KB200_VG_TABLE
    TABLE_NODES = (1 << 20) + 1

    def _device_model(self):
        """model id + stored parameters for the device. Built-in models run as closed forms; a 'custom'
        callable or a GSTools model (which the reference's native 'C' backend refuses,
        variogram_models.pyx:20-21) is tabulated by the host and interpolated on the device
        (KB200_VG_TABLE, include/krige_b200.h: kb200_set_variogram_table)."""
        name = getattr(self.variogram_function, "__name__", None)
        mid = variogram_models.DEVICE_MODEL_IDS.get(name)
        if mid is None or self.variogram_function is not self.variogram_dict.get(self.variogram_model):
            if not callable(self.variogram_function):
                raise NotImplementedError("backend='cuda' needs a built-in variogram model or a callable f(params, d)")
            return self.TABLE_MODEL_ID, []
        return mid, [float(v) for v in self.variogram_model_parameters]

    def _adjusted_corners(self, lo, hi):
        """Corners of the axis-aligned box [lo, hi] (original coordinates) in the adjusted frame."""
        nd = self._ndim
        x, y, z, v, center, Mt = self._data_arrays()
        Mt = np.asarray(Mt, dtype=float).reshape(nd, nd)
        c = np.asarray(center, dtype=float)
        corners = np.array(np.meshgrid(*[[lo[k], hi[k]] for k in range(nd)], indexing="ij")).reshape(nd, -1).T
        return (corners - c) @ Mt.T + c

    def _table_dmax(self, pred_lo=None, pred_hi=None):
        """Upper bound of every distance the device will evaluate: data-data and data-prediction (the
        distance between two boxes is largest at a pair of corners; the affine anisotropy map keeps them
        corners), with head-room for the moving window's local shift gamma(2 d_k)."""
        if getattr(self, "coordinates_type", "euclidean") == "geographic":
            return 360.0
        x, y, z = self._data_arrays()[:3]
        cols = [x, y] + ([z] if self._ndim == 3 else [])
        dlo = [float(np.min(c)) for c in cols]
        dhi = [float(np.max(c)) for c in cols]
        D = self._adjusted_corners(dlo, dhi)
        pts = D if pred_lo is None else np.vstack([D, self._adjusted_corners(pred_lo, pred_hi)])
        span = np.sqrt(((pts[:, None, :] - D[None, :, :]) ** 2).sum(axis=2)).max()
        need = 2.2 * max(float(span), 1e-300)
        have = getattr(self, "_kb_table_dmax", 0.0)
        if need > have:                       # grow geometrically so that moving prediction windows do not re-tabulate
            self._kb_table_dmax = need if have == 0.0 else max(need, 2.0 * have)
        return self._kb_table_dmax

    def _variogram_table(self, dmax):
        """gamma at the sqrt-spaced nodes d_i = dmax (i/(n-1))^2, cached per (callable, parameters, dmax)."""
        key = (id(self.variogram_function), tuple(np.ravel(np.asarray(self.variogram_model_parameters, dtype=float))),
               float(dmax))
        cached = getattr(self, "_kb_table", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        n = self.TABLE_NODES
        d = dmax * (np.arange(n, dtype=np.float64) / (n - 1)) ** 2
        with np.errstate(all="ignore"):
            g = np.asarray(self.variogram_function(self.variogram_model_parameters, d), dtype=np.float64)
        if g.shape != d.shape:
            g = np.broadcast_to(g, d.shape).copy()
        if not np.all(np.isfinite(g)):
            raise ValueError("the custom variogram function must be finite on [0, %g]" % dmax)
        self._kb_table = (key, g)
        return g

    def _data_arrays(self):
        """(x, y, z|None, values, center, Mt) in ORIGINAL coordinates."""
        raise NotImplementedError

    def _drift_spec(self):
        """(n_rl, [host drift data columns])"""
        return 0, []

    def _cuda_handle(self):
        h = getattr(self, "_kb_handle", None)
        if h is None:
            h = _cabi.Handle()
            self._kb_handle = h
            self._kb_key = None
        return h

    def _problem_signature(self, dtype, knn):
        x, y, z, v, center, Mt = self._data_arrays()
        mid, vp = self._device_model()
        n_rl, cols = self._drift_spec()
        if mid == self.TABLE_MODEL_ID:      # the table itself is part of the problem
            vp = ("table", id(self.variogram_function), getattr(self, "_kb_table_dmax", 0.0)) + tuple(
                np.ravel(np.asarray(self.variogram_model_parameters, dtype=float)))
        return (dtype, knn, mid, tuple(vp), bool(self.exact_values), tuple(np.ravel(Mt)), tuple(center),
                n_rl, len(cols), x.size, getattr(self, "coordinates_type", "euclidean"),
                bool(getattr(self, "pseudo_inv", False)))

    def _ensure_problem(self, dtype="float64", knn=False):
        name = dtype if isinstance(dtype, str) and dtype in _cabi.DTYPES else str(np.dtype(dtype))
        dt = _cabi.DTYPES.get(name)
        if dt is None:
            raise ValueError("dtype must be 'float64', 'float32' or 'float64x'")
        h = self._cuda_handle()
        if self._device_model()[0] == self.TABLE_MODEL_ID:
            self._table_dmax()                  # fixes the tabulated range before it enters the signature
        key = self._problem_signature(dt, knn)
        if self._kb_key == key:
            return h
        x, y, z, v, center, Mt = self._data_arrays()
        mid, vp = self._device_model()
        n_rl, cols = self._drift_spec()
        self._kb_key = None
        h.set_coordinates(getattr(self, "coordinates_type", "euclidean") == "geographic")
        h.set_pseudo_inverse(bool(getattr(self, "pseudo_inv", False)) and not knn)
        if mid == self.TABLE_MODEL_ID:
            dmax = self._table_dmax()
            h.set_variogram_table(self._variogram_table(dmax), dmax)
        if knn:
            h.set_problem_knn(self._ndim, x, y, z, v, center, Mt, mid, vp, self.exact_values, self.eps)
        else:
            h.set_problem(self._ndim, dt, x, y, z, v, center, Mt, mid, vp, self.exact_values, self.eps,
                          n_rl=n_rl, drift_data=cols if cols else None)
        self._kb_key = key
        return h

    def _run_cuda(self, style, axes, mask, n_closest_points=None, drift_at=None, dtype="float64"):
        """axes: list of 1-D coordinate arrays [x, y(, z)] (grid axes or point lists, original coords).
        mask: flattened bool mask (True = skip) or None.  drift_at: callable(pts list) -> [n_hd, m]
        host-supplied drift values at the given points, or None.
        Returns flat (z, ss) of length npt in the reference's flattened order."""
        knn = n_closest_points is not None
        nd = self._ndim
        if self._device_model()[0] == self.TABLE_MODEL_ID and all(np.size(a) for a in axes[:nd]):
            self._table_dmax([float(np.min(a)) for a in axes[:nd]], [float(np.max(a)) for a in axes[:nd]])
        h = self._ensure_problem(dtype, knn)
        if style == "points":
            pts = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
            dv = drift_at(pts, None) if drift_at is not None else None
            if knn:
                return h.execute_knn_points(n_closest_points, pts[0], pts[1], pts[2] if nd == 3 else None)
            return h.execute_points(pts[0], pts[1], pts[2] if nd == 3 else None, dv)
        # grid / masked
        gx, gy = axes[0], axes[1]
        gz = axes[2] if nd == 3 else None
        nx, ny = gx.size, gy.size
        nz = gz.size if nd == 3 else 1
        npt = nx * ny * nz
        need_points = (mask is not None) or (drift_at is not None)
        if not need_points:
            if knn:
                return h.execute_knn_grid(n_closest_points, gx, gy, gz)
            return h.execute_grid(gx, gy, gz)
        idx = np.flatnonzero(~mask) if mask is not None else np.arange(npt)
        ix = idx % nx
        iy = (idx // nx) % ny
        pts = [np.asarray(gx, dtype=np.float64)[ix], np.asarray(gy, dtype=np.float64)[iy]]
        if nd == 3:
            pts.append(np.asarray(gz, dtype=np.float64)[idx // (nx * ny)])
        dv = drift_at(pts, idx) if drift_at is not None else None
        if idx.size:
            if knn:
                zc, sc = h.execute_knn_points(n_closest_points, pts[0], pts[1], pts[2] if nd == 3 else None)
            else:
                zc, sc = h.execute_points(pts[0], pts[1], pts[2] if nd == 3 else None, dv)
        else:
            zc = sc = np.zeros(0)
        if mask is None:
            return zc, sc
        z = np.zeros(npt)
        ss = np.zeros(npt)
        z[idx] = zc
        ss[idx] = sc
        return z, ss

    @staticmethod
    def _check_backend(backend, what):
        if backend != "cuda":
            raise ValueError(
                "Specified backend {} is not supported for {}: this package implements backend='cuda' only "
                "(the reference's 'vectorized'/'loop'/'C' CPU paths live in PyKrige).".format(backend, what)
            )
