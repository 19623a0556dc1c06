"""loss/loss_.py mirror (the reference's "chamfer_python": cd, distChamfer, batch_NN_loss, fscore) against golden vectors produced by
the reference's own file on the CPU (tests/golden/make_golden_loss_.py).  CPU part: the oracle reproduces the goldens (distances to
fp32 rounding, indices wherever the reference's float64 expansion has no near-tie); GPU part: the CUDA-backed mirror does."""
import glob
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "loss__*_ref.npz")))


def _close(got, want):
    return np.allclose(got, want, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[7:-8] for p in GOLD])
def test_oracle_reproduces_the_reference_loss_module(oracle, path):
    g = np.load(path)
    d1, d2, i1, i2 = oracle.chamfer_forward(g["x"], g["y"], nthreads=4)
    assert _close(d1, g["d1"]) and _close(d2, g["d2"]) and _close(d1, g["mins2"]) and _close(d2, g["mins1"])
    # the argmin agrees wherever the exact fp32 distances along the two indices are equal or the index is the same
    for idx, ridx, q, t in ((i1, g["i1"], g["x"], g["y"]), (i2, g["i2"], g["y"], g["x"])):
        diff = np.argwhere(idx != ridx)
        for b, j in diff:
            da = ((q[b, j] - t[b, idx[b, j]]).astype(np.float64) ** 2).sum()
            db = ((q[b, j] - t[b, ridx[b, j]]).astype(np.float64) ** 2).sum()
            assert abs(da - db) <= 1e-6 * max(da, db, 1e-12), "a different argmin that is not a tie"
    c1, c2 = oracle.fscore_counts(d1, d2, float(g["thr"]))
    n, m = d1.shape[1], d2.shape[1]
    assert abs((c2 / m).mean() - float(g["p1"])) <= 1.0 / m and abs((c1 / n).mean() - float(g["p2"])) <= 1.0 / n
    assert abs(float(g["nn_loss"]) - (d1.astype(np.float64).mean() + d2.astype(np.float64).mean())) <= 1e-5 * float(g["nn_loss"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[7:-8] for p in GOLD])
def test_cuda_mirror_matches_the_reference_loss_module(pkg, cuda, path):
    g = np.load(path)
    L = pkg.loss_
    x, y = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["y"]).to(cuda)
    loss, mins1, mins2 = L.batch_NN_loss(x, y)
    assert mins1.dtype == torch.float64 and mins1.shape == (x.shape[0], y.shape[1]) and mins2.shape == x.shape[:2]
    assert _close(mins1.cpu().numpy(), g["mins1"]) and _close(mins2.cpu().numpy(), g["mins2"])
    assert abs(float(loss) - float(g["nn_loss"])) <= 1e-5 * float(g["nn_loss"])
    d1, d2, i1, i2 = L.distChamfer(x, y)
    assert d1.dtype == torch.float32 and i1.dtype == torch.int32
    assert _close(d1.cpu().numpy(), g["d1"]) and _close(d2.cpu().numpy(), g["d2"])
    assert (i1.cpu().numpy() != g["i1"]).mean() <= 0.02 and (i2.cpu().numpy() != g["i2"]).mean() <= 0.02 or "lattice" in path or "dup" in path
    f, p1, p2 = L.fscore(x, y, float(g["thr"]))
    n, m = x.shape[1], y.shape[1]
    assert f.requires_grad and abs(float(p1) - float(g["p1"])) <= 1.0 / m and abs(float(p2) - float(g["p2"])) <= 1.0 / n
    assert abs(float(f) - float(g["fscore"])) <= 2.0 / min(n, m)
    assert abs(float(L.cd(x, y)) - float(g["nn_loss"])) <= 1e-5 * float(g["nn_loss"])
    # the dense baseline formula, for completeness (plain torch)
    P = L.batched_pairwise_dist(x[:, :64], y)
    assert P.dtype == torch.float64 and _close(P.min(2)[0].float().cpu().numpy(), g["d1"][:, :64])
