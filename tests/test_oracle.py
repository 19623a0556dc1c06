"""CPU tests of the oracle itself: C restatement vs the independent numpy twin, vs the reference's pure-torch
chamfer (loss/loss_.py, restated), vs committed golden vectors, and domain properties."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import make_clouds

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("kind", ["uniform", "clustered", "lattice", "dup", "offset"])
@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 100, 257), (1, 1, 1), (2, 7, 513), (1, 600, 3)])
def test_chamfer_c_equals_numpy_twin(oracle, kind, shape):
    b, n, m = shape
    x, y = make_clouds(kind, b, n, m, seed=n * 7 + m)
    c = oracle.chamfer_forward(x, y)
    t = oracle.np_chamfer_forward(x, y)
    for a, bb in zip(c, t):
        assert np.array_equal(a, bb)


def test_chamfer_threads_do_not_change_results(oracle):
    x, y = make_clouds("uniform", 5, 300, 400, seed=2)
    a = oracle.chamfer_forward(x, y, nthreads=1)
    b = oracle.chamfer_forward(x, y, nthreads=4)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_chamfer_vs_reference_torch_cpu(oracle):
    """loss_.distChamfer (fp64 expansion) agrees on idx for well-separated random clouds and on dist to 1e-5 rel."""
    x, y = make_clouds("uniform", 4, 512, 640, seed=3)
    c = oracle.chamfer_forward(x, y)
    t = oracle.torch_dist_chamfer(torch.from_numpy(x), torch.from_numpy(y))
    assert np.array_equal(c[2], t[2].numpy()) and np.array_equal(c[3], t[3].numpy())
    assert np.allclose(c[0], t[0].numpy(), rtol=1e-5, atol=1e-9) and np.allclose(c[1], t[1].numpy(), rtol=1e-5, atol=1e-9)


def test_chamfer_lowest_index_ties(oracle):
    y = np.zeros((1, 9, 3), np.float32)
    y[0, :, 0] = [1, 0, 0, 2, 0, 3, 0, 1, 0]     # targets 1,2,4,6,8 coincide at the origin
    x = np.zeros((1, 2, 3), np.float32)
    x[0, 1, 0] = 1                                # equidistant from 0/7 (exact) -> lowest index 0
    d1, d2, i1, i2 = oracle.chamfer_forward(x, y)
    assert i1.tolist() == [[1, 0]] and d1.tolist() == [[0.0, 0.0]]
    assert i2[0].tolist() == [1, 0, 0, 1, 0, 1, 0, 1, 0]


def test_chamfer_nan_tile_semantics(oracle):
    """chamfer3D.cu:36,126: a NaN at the first slot of a 512-tile kills that tile; at tile 0 it poisons the result."""
    x = np.zeros((1, 1, 3), np.float32)
    y = np.ones((1, 1100, 3), np.float32)
    y[0, 600] = 0.25
    y[0, 1050] = 0.0                              # true nearest (d=0) lives in tile 2
    y[0, 1024, 0] = np.nan                        # first element of tile 2 -> tile ignored
    d1, _, i1, _ = oracle.chamfer_forward(x, y)
    assert i1[0, 0] == 600 and d1[0, 0] == np.float32(3 * 0.0625)
    y[0, 0, 0] = np.nan
    d1, _, i1, _ = oracle.chamfer_forward(x, y)
    assert np.isnan(d1[0, 0]) and i1[0, 0] == 0


def test_chamfer_backward_is_the_analytic_gradient(oracle):
    x, y = make_clouds("uniform", 2, 50, 60, seed=4)
    d1, d2, i1, i2 = oracle.chamfer_forward(x, y)
    g1 = np.random.default_rng(0).random((2, 50), dtype=np.float32)
    g2 = np.random.default_rng(1).random((2, 60), dtype=np.float32)
    gx, gy = oracle.chamfer_backward(x, y, g1, g2, i1, i2)
    tx = torch.from_numpy(x).double().requires_grad_(True)
    ty = torch.from_numpy(y).double().requires_grad_(True)
    P = ((tx[:, :, None] - ty[:, None]) ** 2).sum(-1)
    loss = (P.min(2)[0] * torch.from_numpy(g1).double()).sum() + (P.min(1)[0] * torch.from_numpy(g2).double()).sum()
    loss.backward()
    assert np.allclose(gx, tx.grad.numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(gy, ty.grad.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("cfg", [("uniform", 2, 1024, 0.005, 50), ("clustered", 1, 1024, 0.05, 80), ("lattice", 1, 1024, 0.005, 12)])
def test_emd_c_equals_numpy_twin(oracle, cfg):
    kind, b, n, eps, iters = cfg
    x, y = make_clouds(kind, b, n, n, seed=17)
    dc, ac = oracle.emd_forward(x, y, eps, iters)[:2]
    dn, an = oracle.np_emd_forward(x, y, eps, iters)
    assert np.array_equal(ac, an) and np.array_equal(dc, dn)


def test_emd_properties(oracle):
    """Reference-stated invariants (metric/emd/test.py:24-28, README): dist is the squared distance along the
    assignment; every point ends assigned; with enough iterations the assignment becomes a bijection."""
    x, y = make_clouds("uniform", 2, 1024, 1024, seed=23)
    d, a, st = oracle.emd_forward(x, y, 0.05, 1500, nthreads=2, want_stats=True)
    assert a.min() >= 0 and a.max() < 1024
    sel = np.take_along_axis(y, a[..., None].astype(np.int64).repeat(3, -1), 1)
    assert np.allclose(((x - sel) ** 2).sum(-1), d, rtol=1e-5, atol=1e-7)
    assert all(len(set(a[i].tolist())) == 1024 for i in range(2)), "auction did not converge to a bijection"
    u = st["u_hist"]
    assert (u[:, 0] == 1024).all() and (u[:, -1] == 0).all()
    # shape violations (emd_cuda.cu:236-249 / emd_module.py:36-39)
    with pytest.raises(AssertionError):
        oracle.emd_forward(x[:, :1000], y[:, :1000], 0.05, 3)


def test_emd_backward_formula(oracle):
    x, y = make_clouds("uniform", 1, 1024, 1024, seed=29)
    d, a = oracle.emd_forward(x, y, 0.05, 30)[:2]
    g = np.random.default_rng(2).random((1, 1024), dtype=np.float32)
    gx = oracle.emd_backward(x, y, g, a)
    sel = np.take_along_axis(y, a[..., None].astype(np.int64).repeat(3, -1), 1)
    assert np.allclose(gx, 2 * g[..., None] * (x - sel), rtol=1e-6, atol=1e-7)


def test_fscore_matches_reference_torch(oracle):
    x, y = make_clouds("clustered", 3, 400, 500, seed=31)
    d1, d2, _, _ = oracle.chamfer_forward(x, y)
    c1, c2 = oracle.fscore_counts(d1, d2, 1e-4)
    f, p1, p2 = oracle.fscore_from_counts(c1, c2, 400, 500)
    tf, tp1, tp2 = oracle.torch_fscore(torch.from_numpy(x), torch.from_numpy(y), 1e-4)
    assert abs(float(f) - float(tf)) < 1e-6 and abs(float(p1) - float(tp1)) < 1e-6 and abs(float(p2) - float(tp2)) < 1e-6


def test_golden_vectors(oracle):
    """tests/golden/*.npz: inputs + outputs produced by tests/golden/make_golden.py (oracle here, the
    reference's CUDA extensions on the GPU box where the file says source == 'reference_cuda')."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    assert files, "no golden vectors committed"
    for f in files:
        z = np.load(os.path.join(GOLDEN, f))
        if "meta" not in z.files:       # fixtures of the other mirrors (loss_, ...) are checked by their own tests
            continue
        meta = json.loads(str(z["meta"]))
        if meta["op"] == "chamfer":
            got = oracle.chamfer_forward(z["xyz1"], z["xyz2"])
            for g, k in zip(got, ("dist1", "dist2", "idx1", "idx2")):
                assert np.array_equal(g, z[k]), (f, k)
        elif meta["op"] == "emd":
            d, a = oracle.emd_forward(z["xyz1"], z["xyz2"], meta["eps"], meta["iters"])[:2]
            if meta.get("exact", True):
                assert np.array_equal(a, z["assignment"]) and np.array_equal(d, z["dist"]), f
            else:  # reference run with a multi-winner race on some cloud: compare the loss only
                ref = np.sqrt(z["dist"]).mean()
                assert abs(np.sqrt(d).mean() - ref) <= 1e-5 * ref, f
