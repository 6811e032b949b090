"""GSTools interface check (reference: src/pykrige/compat_gstools.py:6-37). GSTools models reach the kriging
classes as 'custom' callables (ok.py:224-239: variogram_function = model.pykrige_vario, anisotropy from the model);
backend='cuda' tabulates such callables on the host and interpolates them on the device (KB200_VG_TABLE)."""


class GSToolsException(Exception):
    """Exception for GSTools."""


def _gstools():
    try:
        import gstools as gs
    except ImportError:
        return None, None
    try:
        version = list(map(int, gs.__version__.split(".")[:2]))
    except Exception:  # noqa: BLE001
        version = None
    return gs, version


def validate_gstools(model):
    """Validate presence and version of GSTools (compat_gstools.py:21-37)."""
    gs, version = _gstools()
    if gs is None:
        raise GSToolsException("GSTools needs to be installed in order to use their CovModel class.")
    if not isinstance(model, gs.CovModel):
        raise GSToolsException("GSTools: given variogram model is not a CovModel instance.")
    if version is not None and version < [1, 3]:
        raise GSToolsException("GSTools: need at least GSTools v1.3.")
    if getattr(model, "latlon", False) and version is not None and version < [1, 4]:
        raise GSToolsException("GSTools: latlon models in PyKrige are only supported from GSTools v1.4.")
