"""Constructor plumbing shared by the two 2-D classes (the reference spells it out twice: ok.py:208-377 and
uk.py:246-394; update_variogram_model ok.py:379-553 and uk.py:630-790). Written once here: option checks, GSTools
hand-over, float64 copies of the data, anisotropy of the data, variogram initialisation, statistics policy."""
import warnings
import numpy as np

from . import core
from .core import _adjust_for_anisotropy, _make_variogram_parameter_list, _initialize_variogram_model

P_INV_TYPES = ("pinv", "pinvh")
GEO_ANISOTROPY_WARNING = "Anisotropy is not compatible with geographic coordinates. Ignoring user set anisotropy."


class Krige2DMixin:
    _ndim = 2
    _prints_coordinates_type = False

    # ---- pieces ---------------------------------------------------------------------------------------------------
    def _gstools_2d(self, variogram_model, variogram_function, check_latlon):
        """Model selection incl. the GSTools route (ok.py:224-239): returns the overrides a CovModel imposes."""

        def _dim_ok(model):
            from .compat_gstools import validate_gstools

            validate_gstools(model)
            if model.field_dim == 3:
                raise ValueError("GSTools: model dim is not 1 or 2")
            if check_latlon and model.latlon and self.coordinates_type == "euclidean":
                raise ValueError("GSTools: latlon models require geographic coordinates")

        return self._select_variogram(variogram_model, variogram_function, _dim_ok)

    def _adjust_data_2d(self):
        self.X_ADJUSTED, self.Y_ADJUSTED = _adjust_for_anisotropy(
            np.vstack((self.X_ORIG, self.Y_ORIG)).T, [self.XCENTER, self.YCENTER],
            [self.anisotropy_scaling], [self.anisotropy_angle]).T

    def _fit_variogram_2d(self, variogram_parameters, nlags, weight):
        vp = _make_variogram_parameter_list(self.variogram_model, variogram_parameters)
        self.lags, self.semivariance, self.variogram_model_parameters = _initialize_variogram_model(
            np.vstack((self.X_ADJUSTED, self.Y_ADJUSTED)).T, self.Z, self.variogram_model, vp,
            self.variogram_function, nlags, weight, self.coordinates_type, lazy=True)
        if self.verbose:
            if self._prints_coordinates_type:           # ok.py:333 (UniversalKriging has no coordinates_type)
                print("Coordinates type: '%s'" % self.coordinates_type, "\n")
            self._print_variogram()
        if self.enable_plotting:
            self.display_variogram_model()

    # ---- constructor body -------------------------------------------------------------------------------------------
    def _init_common_2d(self, x, y, z, variogram_model, variogram_parameters, variogram_function, nlags, weight,
                        anisotropy_scaling, anisotropy_angle, verbose, enable_plotting, exact_values, pseudo_inv,
                        pseudo_inv_type, coordinates_type="euclidean", statistics="lazy"):
        """statistics: 'off' (OrdinaryKriging default), 'eager' (enable_statistics=True) or 'lazy' (UniversalKriging: the
        reference computes them in the constructor, uk.py:380; here on first access)."""
        self.pseudo_inv = bool(pseudo_inv)
        self.pseudo_inv_type = str(pseudo_inv_type)
        if self.pseudo_inv_type not in P_INV_TYPES:
            raise ValueError("pseudo inv type not valid: " + str(pseudo_inv_type))
        if not isinstance(exact_values, bool):
            raise ValueError("exact_values has to be boolean True or False")
        if coordinates_type not in ("euclidean", "geographic"):
            raise ValueError("Only 'euclidean' and 'geographic' are valid values for coordinates-keyword.")
        self.exact_values = exact_values
        self.coordinates_type = coordinates_type
        self.verbose = verbose
        self.enable_plotting = enable_plotting

        ov = self._gstools_2d(variogram_model, variogram_function, check_latlon=True)
        if "gstools" in ov:
            variogram_parameters = []
            anisotropy_scaling, anisotropy_angle = ov["gstools"].pykrige_anis, ov["gstools"].pykrige_angle

        # 1-D float64 copies of the inputs (ok.py:262-268)
        self.X_ORIG, self.Y_ORIG, self.Z = (np.atleast_1d(np.squeeze(np.array(a, copy=True, dtype=np.float64)))
                                            for a in (x, y, z))
        if self.enable_plotting and self.verbose:
            print("Plotting Enabled\n")

        if coordinates_type == "geographic":
            # lon/lat in degrees; anisotropy is ambiguous on the sphere and ignored (ok.py:292-306)
            if anisotropy_scaling != 1.0:
                warnings.warn(GEO_ANISOTROPY_WARNING, UserWarning)
            self.XCENTER = self.YCENTER = 0.0
            self.anisotropy_scaling, self.anisotropy_angle = 1.0, 0.0
            self.X_ADJUSTED, self.Y_ADJUSTED = self.X_ORIG, self.Y_ORIG
        else:
            self.XCENTER = (np.amax(self.X_ORIG) + np.amin(self.X_ORIG)) / 2.0
            self.YCENTER = (np.amax(self.Y_ORIG) + np.amin(self.Y_ORIG)) / 2.0
            self.anisotropy_scaling, self.anisotropy_angle = anisotropy_scaling, anisotropy_angle
            if self.verbose:
                print("Adjusting data for anisotropy...")
            self._adjust_data_2d()

        if self.verbose:
            print("Initializing variogram model...")
        self._fit_variogram_2d(variogram_parameters, nlags, weight)

        self._statistics_policy(statistics)

    def _stats_inputs(self):
        return np.vstack((self.X_ADJUSTED, self.Y_ADJUSTED)).T, self.Z

    def update_variogram_model(self, variogram_model, variogram_parameters=None, variogram_function=None, nlags=6,
                               weight=False, anisotropy_scaling=1.0, anisotropy_angle=0.0):
        """Change the variogram model and/or its parameters (ok.py:379-553, uk.py:630-790). The statistics are
        recomputed on their next access (the reference recomputes them here)."""
        ov = self._gstools_2d(variogram_model, variogram_function, check_latlon=False)
        if "gstools" in ov:
            variogram_parameters = []
            anisotropy_scaling, anisotropy_angle = ov["gstools"].pykrige_anis, ov["gstools"].pykrige_angle
        if self.coordinates_type == "geographic":
            if anisotropy_scaling != 1.0:
                warnings.warn(GEO_ANISOTROPY_WARNING, UserWarning)
        elif (anisotropy_scaling, anisotropy_angle) != (self.anisotropy_scaling, self.anisotropy_angle):
            if self.verbose:
                print("Adjusting data for anisotropy...")
            self.anisotropy_scaling, self.anisotropy_angle = anisotropy_scaling, anisotropy_angle
            self._adjust_data_2d()
        if self.verbose:
            print("Updating variogram mode...")
        self._fit_variogram_2d(variogram_parameters, nlags, weight)
        self._statistics_policy("lazy")

    # ---- device description -------------------------------------------------------------------------------------------
    def _data_arrays(self):
        Mt = core.anisotropy_matrix(2, [self.anisotropy_scaling], [self.anisotropy_angle])
        return self.X_ORIG, self.Y_ORIG, None, self.Z, [self.XCENTER, self.YCENTER], Mt
