"""Farthest point sampling (utils/utils.py:335-360) and the cont_proj splat (utils/projection.py:4-67): numpy oracles pinned
to golden vectors produced by the reference's own torch code on the CPU (tests/golden/make_golden_fps_splat.py); CUDA kernels
(through the C ABI) against the goldens and the oracle.  FPS indices are bit-exact; the splat is float32 with the reference's
rounding sequence, tolerance 2e-6 relative (the last ulps of expf)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from conftest import make_clouds

HERE = os.path.dirname(__file__)
FPS = sorted(glob.glob(os.path.join(HERE, "golden", "fps_*_ref.npz")))
SPLAT = sorted(glob.glob(os.path.join(HERE, "golden", "splat_*_ref.npz")))
SPLAT_RTOL = 2e-6


def ids(paths):
    return [os.path.basename(p) for p in paths]


def test_goldens_exist():
    assert len(FPS) >= 5 and len(SPLAT) >= 4


@pytest.mark.parametrize("path", FPS, ids=ids(FPS))
def test_fps_oracle_matches_reference_python(oracle, path):
    z = np.load(path); meta = json.loads(str(z["meta"]))
    assert np.array_equal(oracle.farthest_point_sample(z["xyz"], meta["npoint"], meta["RAN"]), z["centroids"])


@pytest.mark.parametrize("path", SPLAT, ids=ids(SPLAT))
def test_splat_oracle_matches_reference_python(oracle, path):
    z = np.load(path); meta = json.loads(str(z["meta"]))
    img = oracle.cont_proj(z["pcl"], meta["grid_h"], meta["grid_w"], meta["sigma_sq"])
    assert np.allclose(img, z["image"], rtol=SPLAT_RTOL, atol=1e-30)


def test_fps_oracle_properties(oracle):
    x, _ = make_clouds("uniform", 2, 500, 4, seed=3)
    c = oracle.farthest_point_sample(x, 500, True)
    for b in range(2):
        assert sorted(c[b].tolist()) == list(range(500))       # sampling every point visits each exactly once
    assert (oracle.farthest_point_sample(x, 10, False)[:, 0] == 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FPS, ids=ids(FPS))
def test_fps_kernel_matches_reference_python(pkg, cuda, path):
    z = np.load(path); meta = json.loads(str(z["meta"]))
    got = pkg.utils.farthest_point_sample(torch.from_numpy(z["xyz"]), meta["npoint"], RAN=meta["RAN"])
    assert got.dtype == torch.long and not got.is_cuda          # like the reference: long indices on the input's device
    assert np.array_equal(got.numpy(), z["centroids"])
    pts = pkg.utils.index_points(torch.from_numpy(z["xyz"]), got)
    assert np.array_equal(pts.numpy(), np.take_along_axis(z["xyz"], z["centroids"][..., None].repeat(3, -1), 1))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["uniform", "clustered", "lattice", "dup", "offset"])
@pytest.mark.parametrize("n,npoint", [(1, 1), (2, 2), (33, 20), (512, 128), (513, 256), (2048, 256), (5000, 100), (16384, 64)])
def test_fps_kernel_matches_oracle(pkg, oracle, cuda, kind, n, npoint):
    x, _ = make_clouds(kind, 3, n, 8, seed=n + npoint)
    for ran in (True, False):
        if not ran and n < 2:
            continue
        got = pkg.utils.farthest_point_sample(torch.from_numpy(x).to(cuda), npoint, RAN=ran)
        assert got.is_cuda
        assert np.array_equal(got.cpu().numpy(), oracle.farthest_point_sample(x, npoint, ran))


@pytest.mark.gpu
def test_fps_kernel_special_values_and_errors(pkg, oracle, cuda):
    x, _ = make_clouds("uniform", 2, 300, 8, seed=9)
    x[0, 5] = np.nan; x[1, 7, 0] = np.inf; x[1, 9] = 3e19       # NaN / inf never enter `distance`; 3e19^2 overflows to inf
    got = pkg.utils.farthest_point_sample(torch.from_numpy(x), 50)
    with np.errstate(all="ignore"):
        want = oracle.farthest_point_sample(x, 50, True)
    assert np.array_equal(got.numpy(), want)
    lib = __import__("importlib").import_module(pkg.__name__ + "._lib").lib
    import ctypes
    t = torch.zeros(1, 1, 3, device=cuda); c = torch.zeros(1, 4, dtype=torch.long, device=cuda)
    assert lib.psd_farthest_point_sample(ctypes.c_void_p(t.data_ptr()), 1, 1, 4, 1, ctypes.c_void_p(c.data_ptr()), None) == -1
    big = torch.zeros(1, 20000, 3, device=cuda)
    assert lib.psd_farthest_point_sample(ctypes.c_void_p(big.data_ptr()), 1, 20000, 4, 0, ctypes.c_void_p(c.data_ptr()), None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("path", SPLAT, ids=ids(SPLAT))
def test_splat_kernel_matches_reference_python(pkg, cuda, path):
    z = np.load(path); meta = json.loads(str(z["meta"]))
    img = pkg.projection.cont_proj(torch.from_numpy(z["pcl"]), meta["grid_h"], meta["grid_w"], "cpu", meta["sigma_sq"])
    assert not img.is_cuda and img.shape == z["image"].shape
    assert np.allclose(img.numpy(), z["image"], rtol=SPLAT_RTOL, atol=1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 8, 8), (2, 100, 64, 64), (3, 1024, 64, 64), (1, 513, 65, 130), (2, 64, 1, 7)])
@pytest.mark.parametrize("sigma_sq", [0.5, 0.1, 4.0])
def test_splat_kernel_matches_oracle(pkg, oracle, cuda, shape, sigma_sq):
    b, n, h, w = shape
    rng = np.random.default_rng(n + h)
    pcl = (rng.random((b, n, 3), dtype=np.float32) * 2.4 - 1.2).astype(np.float32)
    img = pkg.projection.cont_proj(torch.from_numpy(pcl).to(cuda), h, w, cuda, sigma_sq)
    want = oracle.cont_proj(pcl, h, w, sigma_sq)
    assert np.allclose(img.cpu().numpy(), want, rtol=SPLAT_RTOL, atol=1e-30)
    # the splat feeds the projection loss: the two ops chain on the device
    mn, mni = pkg.proj_loss.min_dist_terms(img.clamp(0, 1), img.clamp(0, 1),
                                           torch.from_numpy(oracle.grid_dist(h, w).astype(np.float32) + 1))
    assert mn.shape == (b, h, w) and torch.isfinite(mn).all()
