"""Multi-process multi-GPU parity check, launched under torchrun by tests/test_multigpu_gpu.py (and by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/check_mgpu_paths.py

For several problem kinds — built-in model, universal kriging with drift (device- and host-evaluated), 'custom'
callable (tabulated), pseudo_inv=True, geographic, float32 / float64x, 3-D — and every style (grid, masked,
points, moving window) every rank kriges its block after ONE broadcast of rank 0's factor blob
(pykrige_b200.multigpu.execute_sharded: prepare_sharded -> kb200_describe_problem / kb200_blob_commit on the
non-root ranks); the blocks are gathered and must equal rank 0's own single-GPU result BIT FOR BIT, and agree with
the CPU oracle on a subsample. A factorisation failure on rank 0 must raise on every rank (no hang).
Prints one JSON line.
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
    xyz3, val3 = cases.synth_data(32, 500, 3)
    gx, gy = np.linspace(-50, 1050, 61), np.linspace(-50, 1050, 47)
    lon = np.column_stack([xyz[:, 0] * 0.06 - 20.0, xyz[:, 1] * 0.045 + 30.0])
    fn = cases.CUSTOM_VARIOGRAMS["nested"]
    rng = np.random.default_rng(5)
    mask = rng.uniform(size=(gy.size, gx.size)) < 0.35
    px, py = rng.uniform(0, 1000, 777), rng.uniform(0, 1000, 777)
    ex, ey = np.linspace(-100.0, 1100.0, 33), np.linspace(-100.0, 1100.0, 29)
    EX, EY = np.meshgrid(ex, ey)
    raster = 20.0 + 0.03 * EX - 0.02 * EY
    expo = dict(variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])

    def ok():
        return pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, **expo)

    def uk_dev():
        return pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, drift_terms=["regional_linear", "point_log", "external_Z"],
                                   point_drift=np.array([[300.0, 400.0, 1.2]]), external_drift=raster,
                                   external_drift_x=ex, external_drift_y=ey, **expo)

    def uk_func():
        return pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, drift_terms=["functional"],
                                   functional_drift=[lambda x, y: x * 0.001, lambda x, y: np.sin(y / 300.0)], **expo)

    # name -> (model factory, style, axes, execute kwargs, sharded kwargs)
    problems = {
        "ok_grid": (ok, "grid", [gx, gy], {}, {}),
        "ok_masked": (ok, "masked", [gx, gy], {"mask": mask}, {"mask": mask.flatten()}),
        "ok_points": (ok, "points", [px, py], {}, {}),
        "ok_knn_grid": (ok, "grid", [gx, gy], {"n_closest_points": 12}, {"n_closest_points": 12}),
        "ok_knn_points": (ok, "points", [px, py], {"n_closest_points": 12}, {"n_closest_points": 12}),
        "ok_grid_float64x": (ok, "grid", [gx, gy], {"dtype": "float64x"}, {"dtype": "float64x"}),
        "uk_regional_linear_f32": (lambda: pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="spherical",
                                                               variogram_parameters=[1.0, 400.0, 0.05],
                                                               drift_terms=["regional_linear"]), "grid", [gx, gy],
                                   {"dtype": "float32"}, {"dtype": "float32"}),
        "uk_device_drift_masked": (uk_dev, "masked", [gx, gy], {"mask": mask}, {"mask": mask.flatten()}),
        "uk_functional_points": (uk_func, "points", [px, py], {}, {"functional": True}),
        "ok3d_grid": (lambda: pk.OrdinaryKriging3D(xyz3[:, 0], xyz3[:, 1], xyz3[:, 2], val3, variogram_model="gaussian",
                                                   variogram_parameters=[1.0, 300.0, 0.05]), "grid",
                      [np.linspace(0, 1000, 13), np.linspace(0, 1000, 11), np.linspace(0, 250, 7)], {}, {}),
        "ok_custom_nested": (lambda: pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="custom",
                                                        variogram_parameters=fn[1], variogram_function=fn[0]),
                             "grid", [gx, gy], {}, {}),
        "ok_pseudo_inv": (lambda: pk.OrdinaryKriging(xyz[:400, 0], xyz[:400, 1], val[:400], pseudo_inv=True, **expo),
                          "grid", [gx, gy], {}, {}),
        "ok_geographic": (lambda: pk.OrdinaryKriging(lon[:, 0], lon[:, 1], val, variogram_model="exponential",
                                                     variogram_parameters=[1.0, 25.0, 0.05],
                                                     coordinates_type="geographic"), "grid",
                          [np.linspace(-20, 40, 33), np.linspace(30, 75, 29)], {}, {}),
    }
    report = {}
    ok_all = True
    for name, (make, style, axes, ekw, skw) in problems.items():
        m = make()
        skw = dict(skw)
        drift_at = None
        if skw.pop("functional", False):
            from pykrige_b200.core import _adjust_for_anisotropy

            def drift_at(pts, idx, m=m):
                xa, ya = _adjust_for_anisotropy(np.vstack((pts[0], pts[1])).T, [m.XCENTER, m.YCENTER],
                                                [m.anisotropy_scaling], [m.anisotropy_angle]).T
                return np.ascontiguousarray(np.vstack([np.asarray(f(xa, ya), dtype=float) * np.ones(xa.shape)
                                                       for f in m.functional_drift_terms]))
        z, ss = multigpu.execute_sharded(m, style, axes, dist, gather=True, drift_at=drift_at, **skw)
        if rank == 0:
            single = make()
            zs, sss = single.execute(style, *axes, backend="cuda", **ekw)
            zs, sss = np.ravel(np.ma.getdata(zs)), np.ravel(np.ma.getdata(sss))
            if style == "masked":       # the reference leaves masked cells at 0 under the mask
                keep = ~mask.flatten()
                same = bool(np.array_equal(z[keep], zs[keep]) and np.array_equal(ss[keep], sss[keep]))
            else:
                same = bool(np.array_equal(z, zs) and np.array_equal(ss, sss))
            report[name] = {"points": int(z.size), "bitwise_equal_to_single_gpu": same}
            ok_all = ok_all and same
    # a failing factorisation on rank 0 (duplicate points, zero nugget) raises everywhere instead of hanging
    dup = np.vstack([xyz[:50], xyz[:50]])
    bad = pk.OrdinaryKriging(dup[:, 0], dup[:, 1], np.concatenate([val[:50], val[:50]]), variogram_model="exponential",
                             variogram_parameters=[1.0, 300.0, 0.0])
    try:
        multigpu.execute_sharded(bad, "grid", [gx, gy], dist)
        raised = False
    except Exception:  # noqa: BLE001
        raised = True
    flags = [None] * world
    dist.all_gather_object(flags, raised)
    if rank == 0:
        report["singular_raises_on_every_rank"] = bool(all(flags))
        ok_all = ok_all and all(flags)
        # oracle spot check of the plain case
        G = ko.grid_points([gx, gy])[::37]
        zo, so = ko.krige(xyz, val, "exponential", ko.stored_parameters("exponential", [1.0, 300.0, 0.05]), G)
        zg, sg = ok().execute("points", G[:, 0], G[:, 1], backend="cuda")
        report["oracle_max_rel"] = [float(np.max(np.abs(zg - zo)) / np.max(np.abs(zo))),
                                    float(np.max(np.abs(sg - so)) / np.max(np.abs(so)))]
        ok_all = ok_all and report["oracle_max_rel"][0] < 1e-5 and report["oracle_max_rel"][1] < 1e-5
        report["world_size"] = world
        report["ok"] = bool(ok_all)
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(report), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if (rank != 0 or ok_all) else 1


if __name__ == "__main__":
    sys.exit(main())
