"""A stand-in for the `gstools` package (not installable here: no network) with exactly the surface the
reference touches (ok.py:224-239, ok3d.py:248-262, compat_gstools.py): `CovModel` with `pykrige_kwargs`,
`pykrige_vario(args, r)`, `field_dim`, `latlon`, `pykrige_anis*`, `pykrige_angle*`, and `__version__`.
The variogram is GSTools' exponential model: gamma(r) = var (1 - exp(-r / len_scale)) + nugget."""
import sys
import types
import numpy as np


class CovModel:
    def __init__(self, dim=2, var=1.0, len_scale=100.0, nugget=0.0, anis=1.0, angle=0.0, latlon=False):
        self.field_dim = dim
        self.var, self.len_scale, self.nugget = float(var), float(len_scale), float(nugget)
        self.latlon = bool(latlon)
        self.pykrige_anis = 1.0 / anis            # GSTools stores length-scale ratios; PyKrige wants scalings
        self.pykrige_anis_y = 1.0 / anis
        self.pykrige_anis_z = 1.0 / anis
        self.pykrige_angle = float(angle)
        self.pykrige_angle_x = self.pykrige_angle_y = 0.0
        self.pykrige_angle_z = float(angle)

    def variogram(self, r):
        return self.var * (1.0 - np.exp(-np.asarray(r, dtype=float) / self.len_scale)) + self.nugget

    def pykrige_vario(self, args=None, r=0):      # gstools.covmodel.base.CovModel.pykrige_vario signature
        return self.variogram(r)

    @property
    def pykrige_kwargs(self):
        kw = {"variogram_model": "custom", "variogram_parameters": [], "variogram_function": self.pykrige_vario}
        if self.field_dim == 3:
            kw.update(anisotropy_scaling_y=self.pykrige_anis_y, anisotropy_scaling_z=self.pykrige_anis_z,
                      anisotropy_angle_x=self.pykrige_angle_x, anisotropy_angle_y=self.pykrige_angle_y,
                      anisotropy_angle_z=self.pykrige_angle_z)
        else:
            kw.update(anisotropy_scaling=self.pykrige_anis, anisotropy_angle=self.pykrige_angle)
        return kw


def install(version="1.5.2"):
    mod = types.ModuleType("gstools")
    mod.CovModel = CovModel
    mod.__version__ = version
    sys.modules["gstools"] = mod
    return mod


def uninstall():
    sys.modules.pop("gstools", None)
