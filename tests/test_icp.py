"""ICP alignment (utils/icp.py): the numpy oracle is pinned to golden vectors produced by the reference's own icp()
(sklearn KD-tree + numpy SVD, tests/golden/make_golden_icp.py); the batched CUDA kernel (psd_icp_batch / psd_nn_f64 through
the C ABI) is compared with the goldens and with the oracle.  Tolerances: NN indices and iteration counts exact, distances
1e-12 absolute, transforms 1e-9 absolute (fp64 throughout; only summation order and the SVD algorithm differ)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "icp_*_ref.npz")))
IDS = [os.path.basename(p) for p in GOLDEN]
T_TOL, D_TOL = 1e-9, 1e-12


def load(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    pose = z["init_pose"] if z["init_pose"].size else None
    return z, meta, pose


def rot(axis, th):
    axis = np.asarray(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_goldens_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_matches_reference_python(oracle, path):
    z, meta, pose = load(path)
    T, d, i = oracle.icp(z["A"], z["B"], init_pose=pose, max_iterations=meta["max_iterations"], tolerance=meta["tolerance"])
    assert i == int(z["iterations"])
    assert np.array_equal(d, z["distances"])
    assert np.abs(T - z["T"]).max() <= 1e-14
    nd, ni = oracle.icp_nearest_neighbor(z["A"], z["B"])
    assert np.array_equal(ni, z["nn_idx"]) and np.array_equal(nd, z["nn_dist"])
    Tb, _, _ = oracle.icp_best_fit_transform(z["A"].astype(np.float64), z["B"].astype(np.float64))
    assert np.abs(Tb - z["bft_T"]).max() <= 1e-14


def test_oracle_recovers_a_known_rigid_motion(oracle):
    rng = np.random.default_rng(7)
    B = rng.random((200, 3))
    R = rot([1, 2, 3], 0.2)
    A = (B - 0.5) @ R.T + 0.5 + 0.03
    T, d, i = oracle.icp(A, B, max_iterations=200, tolerance=1e-12)
    moved = A @ T[:3, :3].T + T[:3, 3]
    assert np.abs(moved - B).max() < 1e-9 and d.mean() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_kernel_matches_reference_python(pkg, cuda, path):
    z, meta, pose = load(path)
    T, d, i = pkg.icp.icp(z["A"], z["B"], init_pose=pose, max_iterations=meta["max_iterations"], tolerance=meta["tolerance"])
    assert i == int(z["iterations"])
    assert np.abs(d - z["distances"]).max() <= D_TOL
    assert np.abs(T - z["T"]).max() <= T_TOL
    nd, ni = pkg.icp.nearest_neighbor(z["A"], z["B"])
    assert np.array_equal(ni, z["nn_idx"]) and np.array_equal(nd, z["nn_dist"])      # sqrt of the same fp64 sum: bit-equal
    Tb, Rb, tb = pkg.icp.best_fit_transform(z["A"], z["B"])
    assert np.abs(Tb - z["bft_T"]).max() <= T_TOL
    assert np.array_equal(Rb, Tb[:3, :3]) and np.array_equal(tb, Tb[:3, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [64, 512, 513, 1500, 2100, 4096])
def test_kernel_batch_matches_oracle(pkg, oracle, cuda, n, dtype):
    """Every sample of a batch converges on its own (different iteration counts in one launch)."""
    rng = np.random.default_rng(n)
    batch = 3
    A = np.empty((batch, n, 3), dtype); B = np.empty((batch, n, 3), dtype)
    for s in range(batch):
        B[s] = rng.random((n, 3)).astype(dtype)
        R = rot(rng.standard_normal(3), 0.1 + 0.15 * s)
        A[s] = (((B[s].astype(np.float64) - 0.5) @ R.T + 0.5 + 0.02 * s + 0.01 * rng.standard_normal((n, 3))).astype(dtype))[rng.permutation(n)]
    iters = 40 if n > 2000 else 200
    T, d, it = pkg.icp.icp_batch(A, B, max_iterations=iters, tolerance=1e-9)
    T, d, it = T.cpu().numpy(), d.cpu().numpy(), it.cpu().numpy()
    for s in range(batch):
        wT, wd, wi = oracle.icp(A[s], B[s], max_iterations=iters, tolerance=1e-9)
        assert it[s] == wi
        assert np.abs(d[s] - wd).max() <= D_TOL
        assert np.abs(T[s] - wT).max() <= T_TOL
        Rm = T[s][:3, :3]
        assert abs(np.linalg.det(Rm) - 1.0) < 1e-12 and np.abs(Rm @ Rm.T - np.eye(3)).max() < 1e-12


@pytest.mark.gpu
def test_kernel_reflection_and_degenerate_cases(pkg, oracle, cuda):
    rng = np.random.default_rng(11)
    # mirrored correspondences: the unconstrained optimum is a reflection, the rule of utils/icp.py:32-35 must pick a rotation
    B = rng.random((300, 3))
    A = B * np.array([1.0, 1.0, -1.0]) + 0.1
    T, R, t = pkg.icp.best_fit_transform(A, B)
    wT, wR, wt = oracle.icp_best_fit_transform(A, B)
    assert np.linalg.det(R) > 0 and np.abs(T - wT).max() <= T_TOL
    # planar cloud (rank-2 H): still a unique proper rotation
    B2 = rng.random((200, 3)); B2[:, 2] = 0.25
    A2 = (B2 - 0.5) @ rot([0, 0, 1], 0.4).T + 0.5
    T2, _, _ = pkg.icp.best_fit_transform(A2, B2)
    wT2, _, _ = oracle.icp_best_fit_transform(A2, B2)
    assert np.abs(T2 - wT2).max() <= 1e-8
    # identical clouds: zero distances, stops in the first iteration (prev_error starts at 0), identity up to the float32
    # centroid arithmetic of the final best_fit_transform
    B32 = B.astype(np.float32)
    T3, d3, i3 = pkg.icp.icp(B32, B32, max_iterations=10, tolerance=1e-10)
    wT3, wd3, wi3 = oracle.icp(B32, B32, max_iterations=10, tolerance=1e-10)
    assert np.abs(T3 - np.eye(4)).max() < 1e-6 and d3.max() == 0.0 and i3 == wi3 == 0 and np.abs(T3 - wT3).max() <= T_TOL
    # a single point / coincident points: H = 0 -> identity rotation, pure translation
    P = np.full((5, 3), 0.3); Q = np.full((5, 3), 0.7)
    T4, R4, t4 = pkg.icp.best_fit_transform(P, Q)
    assert np.array_equal(R4, np.eye(3)) and np.abs(t4 - 0.4).max() < 1e-15


@pytest.mark.gpu
def test_kernel_argument_errors(pkg, cuda):
    lib = __import__("importlib").import_module(pkg.__name__ + "._lib").lib
    a = torch.zeros(1, 5000, 3, device=cuda)
    T = torch.zeros(1, 4, 4, device=cuda, dtype=torch.float64)
    import ctypes
    rc = lib.psd_icp_batch(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(a.data_ptr()), 0, 1, 5000, None, 5, ctypes.c_double(1e-3),
                           ctypes.c_void_p(T.data_ptr()), None, None, None)
    assert rc == -1                                     # more points than fit in shared memory
    with pytest.raises(AssertionError):
        pkg.icp.icp(np.zeros((4, 3)), np.zeros((5, 3)))  # the reference asserts A.shape == B.shape
