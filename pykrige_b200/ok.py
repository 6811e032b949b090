"""OrdinaryKriging (2-D) with the B200 ``backend='cuda'`` execute() path.

API mirror of the reference class (src/pykrige/ok.py:187-1020): same constructor
arguments, public attributes and ``execute`` signature; ``execute`` runs on the GPU
through libkrige_b200.so instead of scipy (no CPU fallback).
"""
import warnings
import numpy as np

from . import core
from ._base import KrigeBase
from .core import _adjust_for_anisotropy, _make_variogram_parameter_list, _initialize_variogram_model

P_INV_TYPES = ("pinv", "pinvh")


class OrdinaryKriging(KrigeBase):
    """Two-dimensional ordinary kriging; see the reference docstring (ok.py:42-175) for the
    meaning of every argument. Only ``execute(..., backend='cuda')`` differs."""

    _ndim = 2

    def __init__(
        self,
        x,
        y,
        z,
        variogram_model="linear",
        variogram_parameters=None,
        variogram_function=None,
        nlags=6,
        weight=False,
        anisotropy_scaling=1.0,
        anisotropy_angle=0.0,
        verbose=False,
        enable_plotting=False,
        enable_statistics=False,
        coordinates_type="euclidean",
        exact_values=True,
        pseudo_inv=False,
        pseudo_inv_type="pinv",
    ):
        self.pseudo_inv = bool(pseudo_inv)
        self.pseudo_inv_type = str(pseudo_inv_type)
        if self.pseudo_inv_type not in P_INV_TYPES:
            raise ValueError("pseudo inv type not valid: " + str(pseudo_inv_type))
        if not isinstance(exact_values, bool):
            raise ValueError("exact_values has to be boolean True or False")
        self.exact_values = exact_values
        self.coordinates_type = coordinates_type

        def _dim_ok(model):
            from .compat_gstools import validate_gstools

            validate_gstools(model)
            if model.field_dim == 3:
                raise ValueError("GSTools: model dim is not 1 or 2")
            if model.latlon and (self.coordinates_type == "euclidean"):
                raise ValueError("GSTools: latlon models require geographic coordinates")

        ov = self._select_variogram(variogram_model, variogram_function, _dim_ok)
        if "gstools" in ov:
            variogram_parameters = []
            anisotropy_scaling = ov["gstools"].pykrige_anis
            anisotropy_angle = ov["gstools"].pykrige_angle

        # 1-D float64 copies of the inputs (ok.py:262-268)
        self.X_ORIG = np.atleast_1d(np.squeeze(np.array(x, copy=True, dtype=np.float64)))
        self.Y_ORIG = np.atleast_1d(np.squeeze(np.array(y, copy=True, dtype=np.float64)))
        self.Z = np.atleast_1d(np.squeeze(np.array(z, copy=True, dtype=np.float64)))

        self.verbose = verbose
        self.enable_plotting = enable_plotting
        if self.enable_plotting and self.verbose:
            print("Plotting Enabled\n")

        if self.coordinates_type == "euclidean":
            self.XCENTER = (np.amax(self.X_ORIG) + np.amin(self.X_ORIG)) / 2.0
            self.YCENTER = (np.amax(self.Y_ORIG) + np.amin(self.Y_ORIG)) / 2.0
            self.anisotropy_scaling = anisotropy_scaling
            self.anisotropy_angle = anisotropy_angle
            if self.verbose:
                print("Adjusting data for anisotropy...")
            self.X_ADJUSTED, self.Y_ADJUSTED = _adjust_for_anisotropy(
                np.vstack((self.X_ORIG, self.Y_ORIG)).T,
                [self.XCENTER, self.YCENTER],
                [self.anisotropy_scaling],
                [self.anisotropy_angle],
            ).T
        elif self.coordinates_type == "geographic":
            # lon/lat in degrees; anisotropy is ambiguous on the sphere and ignored (ok.py:292-306)
            if anisotropy_scaling != 1.0:
                warnings.warn(
                    "Anisotropy is not compatible with geographic coordinates. Ignoring user set anisotropy.",
                    UserWarning,
                )
            self.XCENTER = 0.0
            self.YCENTER = 0.0
            self.anisotropy_scaling = 1.0
            self.anisotropy_angle = 0.0
            self.X_ADJUSTED = self.X_ORIG
            self.Y_ADJUSTED = self.Y_ORIG
        else:
            raise ValueError("Only 'euclidean' and 'geographic' are valid values for coordinates-keyword.")

        if self.verbose:
            print("Initializing variogram model...")
        vp_temp = _make_variogram_parameter_list(self.variogram_model, variogram_parameters)
        self.lags, self.semivariance, self.variogram_model_parameters = _initialize_variogram_model(
            np.vstack((self.X_ADJUSTED, self.Y_ADJUSTED)).T,
            self.Z,
            self.variogram_model,
            vp_temp,
            self.variogram_function,
            nlags,
            weight,
            self.coordinates_type, lazy=True,
        )
        if self.verbose:
            print("Coordinates type: '%s'" % self.coordinates_type, "\n")
            self._print_variogram()
        if self.enable_plotting:
            self.display_variogram_model()

        if self.verbose:
            print("Calculating statistics on variogram model fit...")
        self._stats_state = "off"
        if enable_statistics:
            self._compute_statistics()
            if self.verbose:
                self.print_statistics()
                print()

    def _stats_inputs(self):
        return np.vstack((self.X_ADJUSTED, self.Y_ADJUSTED)).T, self.Z

    def update_variogram_model(
        self,
        variogram_model,
        variogram_parameters=None,
        variogram_function=None,
        nlags=6,
        weight=False,
        anisotropy_scaling=1.0,
        anisotropy_angle=0.0,
    ):
        """Change the variogram model and/or its parameters (ok.py:379-553)."""

        def _dim_ok(model):
            from .compat_gstools import validate_gstools

            validate_gstools(model)
            if model.field_dim == 3:
                raise ValueError("GSTools: model dim is not 1 or 2")

        ov = self._select_variogram(variogram_model, variogram_function, _dim_ok)
        if "gstools" in ov:
            variogram_parameters = []
            anisotropy_scaling = ov["gstools"].pykrige_anis
            anisotropy_angle = ov["gstools"].pykrige_angle
        if self.coordinates_type == "geographic":
            if anisotropy_scaling != 1.0:
                warnings.warn(
                    "Anisotropy is not compatible with geographic coordinates. Ignoring user set anisotropy.",
                    UserWarning,
                )
        elif anisotropy_scaling != self.anisotropy_scaling or anisotropy_angle != self.anisotropy_angle:
            if self.verbose:
                print("Adjusting data for anisotropy...")
            self.anisotropy_scaling = anisotropy_scaling
            self.anisotropy_angle = anisotropy_angle
            self.X_ADJUSTED, self.Y_ADJUSTED = _adjust_for_anisotropy(
                np.vstack((self.X_ORIG, self.Y_ORIG)).T,
                [self.XCENTER, self.YCENTER],
                [self.anisotropy_scaling],
                [self.anisotropy_angle],
            ).T
        if self.verbose:
            print("Updating variogram mode...")
        vp_temp = _make_variogram_parameter_list(self.variogram_model, variogram_parameters)
        self.lags, self.semivariance, self.variogram_model_parameters = _initialize_variogram_model(
            np.vstack((self.X_ADJUSTED, self.Y_ADJUSTED)).T,
            self.Z,
            self.variogram_model,
            vp_temp,
            self.variogram_function,
            nlags,
            weight,
            self.coordinates_type, lazy=True,
        )
        if self.verbose:
            self._print_variogram()
        if self.enable_plotting:
            self.display_variogram_model()
        # the reference recomputes the statistics eagerly here (ok.py:533-553); lazy instead
        self._stats_state = "lazy"

    # ---- device description -----------------------------------------------------------
    def _data_arrays(self):
        Mt = core.anisotropy_matrix(2, [self.anisotropy_scaling], [self.anisotropy_angle])
        return self.X_ORIG, self.Y_ORIG, None, self.Z, [self.XCENTER, self.YCENTER], Mt

    def execute(self, style, xpoints, ypoints, mask=None, backend="cuda", n_closest_points=None, dtype="float64",
                n_gpus=None):
        """Calculates a kriged grid and the associated variance (ok.py:760-1020).

        ``backend='cuda'`` is the only backend of this package. ``style``, ``mask`` and
        ``n_closest_points`` behave as in the reference, including the exception types.
        ``n_gpus=G`` shards the prediction points over G GPUs of this box from this one host thread.
        Returns ``(zvalues, sigmasq)`` shaped ``(ny, nx)`` for 'grid'/'masked' (masked arrays
        for 'masked') or ``(n,)`` for 'points'.
        """
        if self.verbose:
            print("Executing Ordinary Kriging...\n")
        if style != "grid" and style != "masked" and style != "points":
            raise ValueError("style argument must be 'grid', 'points', or 'masked'")
        if n_closest_points is not None and n_closest_points <= 1:
            raise ValueError("n_closest_points has to be at least two!")
        axes, sizes, flat_mask = self._prepare_points(style, (xpoints, ypoints), mask)
        self._check_backend(backend, "2D ordinary kriging")
        zvalues, sigmasq = self._run_cuda(style, axes, flat_mask, n_closest_points=n_closest_points, dtype=dtype,
                                          n_gpus=n_gpus)
        return self._shape_output(style, zvalues, sigmasq, sizes, flat_mask)
