"""world_size-2 gloo test (CPU) of the N>1 host path: contiguous point sharding + the single
broadcast of the factor data, with the oracle standing in for the device executor. Verifies that
concatenating the rank-local slices reproduces the single-process result exactly."""
import os
import socket
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.linalg
    from scipy.spatial.distance import cdist
    import cases
    from oracle import krige_oracle as ko
    from pykrige_b200 import multigpu

    xyz, val = cases.synth_data(42, 120, 2)
    gx, gy = np.linspace(0, 1000, 17), np.linspace(0, 1000, 13)
    stored = ko.stored_parameters("exponential", [1.0, 300.0, 0.05])
    n = xyz.shape[0]
    # rank 0 "factors"; ONE broadcast ships the factor (here: the dense inverse) to everyone
    blob = torch.zeros((n + 1) * (n + 1), dtype=torch.float64)
    if rank == 0:
        blob.copy_(torch.from_numpy(scipy.linalg.inv(ko.kriging_matrix(xyz, "exponential", stored)).ravel()))
    dist.broadcast(blob, src=0)
    a_inv = blob.numpy().reshape(n + 1, n + 1)
    G = ko.grid_points([gx, gy])
    first, count = multigpu.shard_range(G.shape[0], rank, world)
    Q = G[first:first + count]
    bd = cdist(Q, xyz)
    b = np.ones((count, n + 1))
    b[:, :n] = -ko.variogram("exponential", stored, bd)
    x = a_inv @ b.T
    z = x[:n].T @ val
    ss = -np.einsum("ij,ji->i", b, x)
    np.savez(os.path.join(outdir, "r%d.npz" % rank), z=z, ss=ss, first=first, count=count)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from oracle import krige_oracle as ko
    xyz, val = cases.synth_data(42, 120, 2)
    gx, gy = np.linspace(0, 1000, 17), np.linspace(0, 1000, 13)
    G = ko.grid_points([gx, gy])
    z1, s1 = ko.krige(xyz, val, "exponential", ko.stored_parameters("exponential", [1.0, 300.0, 0.05]), G)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    assert parts[0]["first"] == 0 and parts[0]["count"] + parts[1]["count"] == G.shape[0]
    assert parts[1]["first"] == parts[0]["count"]
    z = np.concatenate([p["z"] for p in parts])
    ss = np.concatenate([p["ss"] for p in parts])
    np.testing.assert_allclose(z, z1, rtol=1e-10)
    np.testing.assert_allclose(ss, s1, rtol=1e-9, atol=1e-12)
