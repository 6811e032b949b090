"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/check_mgpu_paths.py

For several problem kinds — built-in model, universal kriging with drift, 'custom' callable (tabulated),
pseudo_inv=True, geographic, float32 — every rank kriges its shard of the grid after ONE broadcast of rank
0's factor blob (pykrige_b200.multigpu.execute_grid_sharded); the shards are gathered and must equal rank
0's own single-GPU result BIT FOR BIT, and agree with the CPU oracle on a subsample. Prints one JSON line.
"""
import json
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import pykrige_b200 as pk  # noqa: E402
from pykrige_b200 import multigpu  # noqa: E402
from oracle import krige_oracle as ko  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    saved = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xyz, val = cases.synth_data(31, 900, 2)
    gx, gy = np.linspace(-50, 1050, 61), np.linspace(-50, 1050, 47)
    lon = np.column_stack([xyz[:, 0] * 0.06 - 20.0, xyz[:, 1] * 0.045 + 30.0])
    fn = cases.CUSTOM_VARIOGRAMS["nested"]
    problems = {
        "ok_exponential": (lambda: pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential",
                                                      variogram_parameters=[1.0, 300.0, 0.05]), "float64", [gx, gy]),
        "uk_regional_linear_f32": (lambda: pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="spherical",
                                                               variogram_parameters=[1.0, 400.0, 0.05],
                                                               drift_terms=["regional_linear"]), "float32", [gx, gy]),
        "ok_custom_nested": (lambda: pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="custom",
                                                        variogram_parameters=fn[1], variogram_function=fn[0]),
                             "float64", [gx, gy]),
        "ok_pseudo_inv": (lambda: pk.OrdinaryKriging(xyz[:400, 0], xyz[:400, 1], val[:400], variogram_model="exponential",
                                                     variogram_parameters=[1.0, 300.0, 0.05], pseudo_inv=True),
                          "float64", [gx, gy]),
        "ok_geographic": (lambda: pk.OrdinaryKriging(lon[:, 0], lon[:, 1], val, variogram_model="exponential",
                                                     variogram_parameters=[1.0, 25.0, 0.05],
                                                     coordinates_type="geographic"), "float64",
                          [np.linspace(-20, 40, 33), np.linspace(30, 75, 29)]),
    }
    report = {}
    ok_all = True
    for name, (make, dtype, axes) in problems.items():
        m = make()
        z, ss, first, count = multigpu.execute_grid_sharded(m, axes, dist, dtype=dtype)
        parts = [None] * world
        dist.all_gather_object(parts, (first, z, ss))
        if rank == 0:
            parts.sort(key=lambda p: p[0])
            zc = np.concatenate([p[1] for p in parts])
            sc = np.concatenate([p[2] for p in parts])
            single = make()
            zs, sss = single.execute("grid", *axes, backend="cuda", dtype=dtype)
            same = bool(np.array_equal(zc, np.ravel(zs)) and np.array_equal(sc, np.ravel(sss)))
            report[name] = {"points": int(zc.size), "bitwise_equal_to_single_gpu": same}
            ok_all = ok_all and same
    if rank == 0:
        # oracle spot check of the plain case
        G = ko.grid_points([gx, gy])[::37]
        zo, so = ko.krige(xyz, val, "exponential", ko.stored_parameters("exponential", [1.0, 300.0, 0.05]), G)
        m = problems["ok_exponential"][0]()
        zg, sg = m.execute("points", G[:, 0], G[:, 1], backend="cuda")
        report["oracle_max_rel"] = [float(np.max(np.abs(zg - zo)) / np.max(np.abs(zo))),
                                    float(np.max(np.abs(sg - so)) / np.max(np.abs(so)))]
        ok_all = ok_all and report["oracle_max_rel"][0] < 1e-5 and report["oracle_max_rel"][1] < 1e-5
        report["world_size"] = world
        report["ok"] = bool(ok_all)
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
