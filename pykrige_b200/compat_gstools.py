"""GSTools interface check (reference: src/pykrige/compat_gstools.py:21-36). GSTools models reach
the kriging classes as 'custom' callables; backend='cuda' tabulates such callables on the host and
interpolates them on the device (KB200_VG_TABLE)."""


class GSToolsException(Exception):
    pass


def validate_gstools(model):
    try:
        import gstools as gs
    except ImportError:
        raise GSToolsException("GSTools: if you want to use GSTools models, install gstools")
    if not isinstance(model, gs.CovModel):
        raise GSToolsException("GSTools: given variogram model is not a CovModel")
