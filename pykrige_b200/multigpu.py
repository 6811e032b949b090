"""Grid-point sharding across the GPUs of one box.

One process per GPU (torchrun). Every prediction point is independent given the factorisation
(ok.py:679-681 is column-wise independent), so the flattened point index is cut into contiguous
blocks, one per rank. Rank 0 assembles and factors; ONE broadcast (NCCL over NVLink) ships the
factor blob — packed inverse Cholesky factor, dual rows, drift constants, adjusted data
coordinates; no other collective is on the data path. torch is used only as the owner of the
process group and as a zero-copy view of the blob's device memory.
"""
import numpy as np


def shard_range(count, rank, world):
    """Contiguous block [first, first+n) of `count` items for `rank` of `world` (sizes differ by <= 1)."""
    count, rank, world = int(count), int(rank), int(world)
    base, rem = divmod(count, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class _DevicePtr:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {
            "shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2,
        }


def blob_as_tensor(handle, device):
    import torch

    ptr, nbytes = handle.blob()
    if not ptr or not nbytes:
        raise RuntimeError("the handle has no factor blob (describe/set the problem first)")
    return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)


def prepare_sharded(model, dist=None, src=0, dtype="float64"):
    """Make `model` ready to execute on every rank: rank `src` factors, everyone else only describes
    the problem (allocating the blob) and receives the broadcast. Returns the model's C-ABI handle."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return model._ensure_problem(dtype)
    import torch

    rank = dist.get_rank()
    if rank == src:
        h = model._ensure_problem(dtype)
    else:
        h = model._cuda_handle()
        x, y, z, v, center, Mt = model._data_arrays()
        mid, vp = model._device_model()
        n_rl, cols = model._drift_spec()
        from . import _cabi
        dt = _cabi.DTYPES[dtype if dtype in _cabi.DTYPES else str(np.dtype(dtype))]
        h.set_coordinates(getattr(model, "coordinates_type", "euclidean") == "geographic")
        h.set_pseudo_inverse(bool(getattr(model, "pseudo_inv", False)))
        if mid == model.TABLE_MODEL_ID:      # 'custom' callable: every rank tabulates it itself (no broadcast)
            dmax = model._table_dmax()
            h.set_variogram_table(model._variogram_table(dmax), dmax)
        h.describe_problem(model._ndim, dt, x, y, z, v, center, Mt, mid, vp, model.exact_values, model.eps,
                           n_rl=n_rl, drift_data=cols if cols else None)
        model._kb_key = None
    t = blob_as_tensor(h, torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(t, src=src)           # the single collective of the path
    torch.cuda.current_stream().synchronize()
    if rank != src:
        h.blob_commit()
        from . import _cabi as _c
        model._kb_key = model._problem_signature(_c.DTYPES[dtype if dtype in _c.DTYPES else str(np.dtype(dtype))], False)
    return h


def execute_grid_sharded(model, axes, dist=None, dtype="float64"):
    """Krige this rank's contiguous slice of the flattened grid. Returns (z, ss, first, count) with host
    arrays of the slice; concatenating the slices in rank order reproduces the single-GPU result bit
    for bit (per-point arithmetic does not depend on the sharding)."""
    if model._device_model()[0] == model.TABLE_MODEL_ID:
        nd = model._ndim                     # the tabulated range must cover the prediction grid (same on every rank)
        model._table_dmax([float(np.min(a)) for a in axes[:nd]], [float(np.max(a)) for a in axes[:nd]])
    h = prepare_sharded(model, dist, dtype=dtype)
    gx, gy = axes[0], axes[1]
    gz = axes[2] if len(axes) > 2 else None
    npt = len(gx) * len(gy) * (len(gz) if gz is not None else 1)
    if dist is not None and dist.is_initialized():
        first, count = shard_range(npt, dist.get_rank(), dist.get_world_size())
    else:
        first, count = 0, npt
    z, ss = h.execute_grid(gx, gy, gz, None, first, count)
    return z, ss, first, count
