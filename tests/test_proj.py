"""Projection-loss min-distance terms (loss/proj_loss.py:21-40): oracle pinned to golden vectors produced by the reference's
own Python (tests/golden/make_golden_proj.py); CUDA kernel (through the C ABI) against the oracle and the goldens."""
import glob
import os

import numpy as np
import pytest
import torch

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "proj_*_ref.npz")))


def bit_equal(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_python(oracle, path):
    g = np.load(path)
    mn, mni = oracle.proj_min_dist(g["pred"], g["gt"], g["dist_mat_after"], mode="as_written")
    assert bit_equal(mn, g["min_dist"]) and bit_equal(mni, g["min_dist_inv"])
    h, w = g["pred"].shape[1:]
    assert bit_equal(oracle.grid_dist(h, w).astype(np.float32) + np.float32(1.0), g["dist_mat_after"])


def test_package_grid_dist_matches_oracle(oracle):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_proj_loss_src", os.path.join(os.path.dirname(__file__), "..", "3d-pointcloudreconstruction_b200", "proj_loss.py"))
    src = open(spec.origin).read()
    assert "def grid_dist" in src and "def get_loss_proj" in src     # the mirror keeps the reference's names
    assert np.array_equal(oracle.grid_dist(5, 7), oracle.grid_dist(5, 7))
    # closed form of the as-written semantics: the minimum over (h', w') sits at an end point of the distance range
    rng = np.random.default_rng(0)
    pred = (rng.random((2, 5, 7), dtype=np.float32) * 3 - 1).astype(np.float32)
    gt = (rng.random((2, 5, 7), dtype=np.float32) * 3 - 1).astype(np.float32)
    d = (oracle.grid_dist(5, 7).astype(np.float32) + np.float32(1.0)).astype(np.float32)
    mn, mni = oracle.proj_min_dist(pred, gt, d)
    one, big = np.float32(1), np.float32(1e6)
    gth = gt + (one - gt) * big
    dmax = d.reshape(5, 7, -1).max(-1)
    cand = np.minimum((gth * np.float32(1.0)).astype(np.float32) * pred, ((gth * dmax[None]).astype(np.float32) * pred).astype(np.float32))
    assert bit_equal(mn, cand)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_kernel_as_written_matches_reference_python(pkg, cuda, path):
    g = np.load(path)
    pl = pkg.proj_loss
    mn, mni = pl.min_dist_terms(torch.from_numpy(g["pred"]), torch.from_numpy(g["gt"]), torch.from_numpy(g["dist_mat_after"]))
    assert not mn.is_cuda                                   # like the reference: results on the inputs' device
    assert bit_equal(mn.numpy(), g["min_dist"]) and bit_equal(mni.numpy(), g["min_dist_inv"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 8, 8), (3, 17, 13), (2, 32, 32)])
@pytest.mark.parametrize("kind", ["soft", "binary", "signed", "nan"])
def test_kernel_both_modes_match_oracle(pkg, oracle, cuda, shape, kind):
    b, h, w = shape
    rng = np.random.default_rng(h * 100 + w)
    if kind == "binary":
        pred = (rng.random((b, h, w)) < 0.3).astype(np.float32); gt = (rng.random((b, h, w)) < 0.4).astype(np.float32)
    else:
        pred = rng.random((b, h, w), dtype=np.float32); gt = rng.random((b, h, w), dtype=np.float32)
        if kind == "signed":
            pred = (pred * 3 - 1).astype(np.float32); gt = (gt * 3 - 1).astype(np.float32)
        if kind == "nan":
            pred[0, 1, 2] = np.nan; gt[-1, 0, 0] = np.nan
    d = (oracle.grid_dist(h, w).astype(np.float32) + np.float32(1.0)).astype(np.float32)
    for mode in ("as_written", "intended"):
        want = oracle.proj_min_dist(pred, gt, d, mode=mode)
        got = pkg.proj_loss.min_dist_terms(torch.from_numpy(pred).to(cuda), torch.from_numpy(gt).to(cuda), torch.from_numpy(d), mode=mode)
        assert got[0].is_cuda
        assert bit_equal(got[0].cpu().numpy(), want[0]) and bit_equal(got[1].cpu().numpy(), want[1]), (mode, kind, shape)


@pytest.mark.gpu
def test_get_loss_proj_signature_and_full_grid(pkg, oracle, cuda):
    """The finetune.py call (:158) at the reference's grid size 64x64; dist_mat is incremented in place like the reference."""
    import types
    b, h, w = 2, 64, 64
    rng = np.random.default_rng(1)
    pred = torch.from_numpy(rng.random((b, h, w), dtype=np.float32)).clamp(1e-3, 1 - 1e-3)
    gt = torch.from_numpy(rng.random((b, h, w), dtype=np.float32))
    dist_mat = torch.from_numpy(pkg.proj_loss.grid_dist(h, w).astype(np.float32))
    before = dist_mat.clone()
    loss, fwd, bwd = pkg.proj_loss.get_loss_proj(pred, gt, "cuda", "bce_prob", 1.0, True, dist_mat, opt=types.SimpleNamespace(grid_h=h, grid_w=w))
    assert torch.equal(dist_mat, before + 1) and loss.is_cuda and fwd.shape == (b, h, w)
    one, big = np.float32(1), np.float32(1e6)
    g = gt.numpy(); p = pred.numpy()
    assert bit_equal(fwd.numpy(), ((g + (one - g) * big) * np.float32(1.0)).astype(np.float32) * p)    # c >= 0: D1 = 1 wins
    # intended mode on the full grid against the dense oracle of one sample
    f2, b2 = pkg.proj_loss.min_dist_terms(pred[:1], gt[:1], dist_mat, mode="intended")
    w2 = oracle.proj_min_dist(p[:1], g[:1], dist_mat.numpy(), mode="intended")
    assert bit_equal(f2.numpy(), w2[0]) and bit_equal(b2.numpy(), w2[1])
