"""Multi-GPU parity on the GPU box (skipped when fewer than two devices are visible):
  * single-process: execute(..., backend='cuda', n_gpus=G) (kb200_group_*: one host thread, device 0 factors, peer
    copy of the blob, contiguous blocks) must equal the single-GPU result bit for bit for every style;
  * multi-process: tests/check_mgpu_paths.py under torchrun (one rank per GPU, NCCL broadcast of the blob) for all
    styles and problem kinds."""
import json
import os
import socket
import subprocess
import sys
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


needs2 = pytest.mark.skipif(_n_devices() < 2, reason="needs >= 2 CUDA devices")


@pytest.fixture(scope="module")
def pk():
    import pykrige_b200
    return pykrige_b200


def _same(a, b):
    return np.array_equal(np.ma.getdata(a), np.ma.getdata(b))


@needs2
def test_single_process_n_gpus_equals_one_gpu(pk):
    G = min(_n_devices(), 8)
    xyz, val = cases.synth_data(61, 700, 2)
    gx, gy = np.linspace(-20, 1020, 83), np.linspace(-20, 1020, 59)
    rng = np.random.default_rng(9)
    mask = rng.uniform(size=(gy.size, gx.size)) < 0.3
    px, py = rng.uniform(0, 1000, 1001), rng.uniform(0, 1000, 1001)
    kw = dict(variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
    ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, **kw)
    for g in sorted({2, G}):
        for style, axes, ekw in (("grid", (gx, gy), {}), ("masked", (gx, gy), {"mask": mask}), ("points", (px, py), {}),
                                 ("grid", (gx, gy), {"n_closest_points": 10}),
                                 ("points", (px, py), {"n_closest_points": 10}),
                                 ("grid", (gx, gy), {"dtype": "float32"}), ("grid", (gx, gy), {"dtype": "float64x"})):
            z1, s1 = ok.execute(style, *axes, backend="cuda", **ekw)
            zg, sg = ok.execute(style, *axes, backend="cuda", n_gpus=g, **ekw)
            assert z1.shape == zg.shape
            assert _same(z1, zg) and _same(s1, sg), (g, style, ekw)
    # universal kriging: device-evaluated and host-supplied drift columns
    ex, ey = np.linspace(-100.0, 1100.0, 33), np.linspace(-100.0, 1100.0, 29)
    EX, EY = np.meshgrid(ex, ey)
    uk = pk.UniversalKriging(xyz[:, 0], xyz[:, 1], val, drift_terms=["regional_linear", "point_log", "external_Z", "functional"],
                             point_drift=np.array([[300.0, 400.0, 1.2]]), external_drift=20.0 + 0.03 * EX - 0.02 * EY,
                             external_drift_x=ex, external_drift_y=ey,
                             functional_drift=[lambda x, y: np.sin(x / 200.0)], **kw)
    for style, axes, ekw in (("grid", (gx, gy), {}), ("masked", (gx, gy), {"mask": mask}), ("points", (px, py), {})):
        z1, s1 = uk.execute(style, *axes, backend="cuda", **ekw)
        zg, sg = uk.execute(style, *axes, backend="cuda", n_gpus=G, **ekw)
        assert _same(z1, zg) and _same(s1, sg), (style,)
    # 3-D
    xyz3, val3 = cases.synth_data(62, 400, 3)
    k3 = pk.OrdinaryKriging3D(xyz3[:, 0], xyz3[:, 1], xyz3[:, 2], val3, variogram_model="gaussian",
                              variogram_parameters=[1.0, 300.0, 0.05])
    a3 = (np.linspace(0, 1000, 17), np.linspace(0, 1000, 13), np.linspace(0, 250, 9))
    z1, s1 = k3.execute("grid", *a3, backend="cuda")
    zg, sg = k3.execute("grid", *a3, backend="cuda", n_gpus=G)
    assert _same(z1, zg) and _same(s1, sg)
    # more GPUs than the box has -> ValueError; a singular system raises like the single-GPU path
    with pytest.raises(ValueError):
        ok.execute("grid", gx, gy, backend="cuda", n_gpus=_n_devices() + 1)
    dup = np.vstack([xyz[:40], xyz[:40]])
    bad = pk.OrdinaryKriging(dup[:, 0], dup[:, 1], np.concatenate([val[:40], val[:40]]), variogram_model="exponential",
                             variogram_parameters=[1.0, 300.0, 0.0])
    with pytest.raises(np.linalg.LinAlgError):
        bad.execute("grid", gx, gy, backend="cuda", n_gpus=2)


@needs2
def test_single_process_n_gpus_full_size_cfg2(pk):
    """BASELINE config 2 through execute(n_gpus=G): equals the 1-GPU result on a slab, bit for bit."""
    G = min(_n_devices(), 8)
    xyz, val = cases.synth_data(1002, 5000, 2)
    ok = pk.OrdinaryKriging(xyz[:, 0], xyz[:, 1], val, variogram_model="exponential", variogram_parameters=[1.0, 300.0, 0.05])
    gx, gy = np.linspace(0, 1000, 1000), np.linspace(0, 1000, 1000)[:64]
    z1, s1 = ok.execute("grid", gx, gy, backend="cuda")
    zg, sg = ok.execute("grid", gx, gy, backend="cuda", n_gpus=G)
    assert np.array_equal(z1, zg) and np.array_equal(s1, sg)


@needs2
def test_torchrun_sharded_paths():
    """tests/check_mgpu_paths.py under torch.distributed.run with 2 ranks: prepare_sharded / describe_problem /
    blob_commit + every style, gathered shards == single GPU bit for bit."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "check_mgpu_paths.py")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    rep = json.loads(lines[-1])
    assert rep["ok"], rep
    assert rep["world_size"] == 2
    for name, r in rep.items():
        if isinstance(r, dict):
            assert r["bitwise_equal_to_single_gpu"], name
