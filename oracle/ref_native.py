"""Thin driver for the reference's compiled native twins in oracle/_ref (built by oracle/build_ref.py from
/root/reference/src/pykrige/lib/*.pyx) — TEST INFRASTRUCTURE ONLY, same rule as krige_oracle.py.

The compiled functions are the reference's own code; this file only prepares their arguments exactly as
the reference's execute() does for backend='C' (ok.py:887-986): the (n+1)^2 kriging matrix, the M x N
distance matrix, the int8 mask and the `pars` dict (ok.py:916-927).
"""
import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.distance import cdist

from . import build_ref
from . import krige_oracle as ko


class _Named:
    """Stands in for the variogram callable: the Cython table is keyed on __name__ only
    (cok.pyx:37, variogram_models.pyx:8-19)."""

    def __init__(self, model):
        self.__name__ = {"linear": "linear_variogram_model", "power": "power_variogram_model",
                         "gaussian": "gaussian_variogram_model", "exponential": "exponential_variogram_model",
                         "spherical": "spherical_variogram_model"}[model]


def available():
    return build_ref.load() is not None


def _pars(values, model, stored, exact_values):
    return {"Z": np.ascontiguousarray(values, dtype=np.float64), "eps": ko.EPS,
            "variogram_model_parameters": np.asarray(stored, dtype=np.float64),
            "variogram_function": _Named(model), "exact_values": bool(exact_values),
            "pseudo_inv": False, "pseudo_inv_type": "pinv"}


def exec_loop(P, Q, values, model, stored, exact_values=True):
    """_c_exec_loop (cok.pyx:14-96) on adjusted data P [n, dim] and points Q [m, dim]: ordinary kriging,
    global system. Returns (z, ss)."""
    cok = build_ref.load()
    a = np.ascontiguousarray(ko.kriging_matrix(P, model, stored))
    bd = np.ascontiguousarray(cdist(Q, P, "euclidean"))
    mask = np.zeros(Q.shape[0], dtype="int8")
    z, ss = cok._c_exec_loop(a, bd, mask, P.shape[0], _pars(values, model, stored, exact_values))
    return np.asarray(z), np.asarray(ss)


def exec_loop_moving_window(P, Q, values, model, stored, k, exact_values=True):
    """_c_exec_loop_moving_window (cok.pyx:98-193) with the kd-tree query of ok.py:957-960."""
    cok = build_ref.load()
    a = np.ascontiguousarray(ko.kriging_matrix(P, model, stored))
    bd, idx = cKDTree(P).query(Q, k=k, eps=0.0)
    mask = np.zeros(Q.shape[0], dtype="int8")
    z, ss = cok._c_exec_loop_moving_window(a, np.ascontiguousarray(bd), mask,
                                           np.ascontiguousarray(idx, dtype=np.int64), k,
                                           _pars(values, model, stored, exact_values))
    return np.asarray(z), np.asarray(ss)
