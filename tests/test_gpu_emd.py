"""GPU parity: persistent cluster auction (through the C ABI) vs the CPU oracle, bit-exact assignment and dist."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


def run_cluster(pkg, dev, x, y, eps, iters, cluster):
    b, n, _ = x.shape
    tx, ty = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    dist = torch.zeros(b, n, device=dev)
    ass = torch.full((b, n), -1, device=dev, dtype=torch.int32)
    inv = torch.full((b, n), -1, device=dev, dtype=torch.int32)
    price = torch.zeros(b, n, device=dev)
    L = pkg._lib
    rc = L.lib.psd_emd_forward_cluster(L.ptr(tx), L.ptr(ty), b, n, L.ptr(dist), L.ptr(ass), L.ptr(price), L.ptr(inv),
                                       ctypes.c_float(eps), iters, cluster, L.stream_of(tx))
    assert rc == 1, L.last_error()
    torch.cuda.synchronize()
    return dist.cpu().numpy(), ass.cpu().numpy(), price.cpu().numpy(), inv.cpu().numpy()


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("cfg", [("uniform", 3, 1024, 0.005, 50), ("clustered", 2, 2048, 0.005, 50), ("uniform", 2, 1024, 0.05, 300),
                                 ("lattice", 2, 1024, 0.005, 20), ("uniform", 1, 2048, 0.002, 7), ("uniform", 2, 1024, 0.005, 1),
                                 ("dup", 2, 1024, 0.005, 30), ("offset", 2, 1024, 0.005, 30), ("uniform", 1, 4096, 0.005, 12)])
def test_emd_bit_exact_all_cluster_sizes(pkg, oracle, cuda, cluster, cfg):
    kind, b, n, eps, iters = cfg
    x, y = make_clouds(kind, b, n, n, seed=100 + n + iters)
    wd, wa, state = oracle.emd_forward(x, y, eps, iters, nthreads=8, full_state=True)
    import os
    # with the object grid (default threshold: only iterations with many bidders), with the grid in EVERY iteration
    # (PSD_EMD_GRID_MIN_U=0: exercises the second shell and the every-object fallback), and without it
    for grid, min_u in ((1, None), (1, "0"), (0, None)):
        oldg = pkg._lib.lib.psd_emd_grid_mode(grid)
        if min_u is not None:
            os.environ["PSD_EMD_GRID_MIN_U"] = min_u
        try:
            dist, ass, price, inv = run_cluster(pkg, cuda, x, y, eps, iters, cluster)
        finally:
            pkg._lib.lib.psd_emd_grid_mode(oldg)
            os.environ.pop("PSD_EMD_GRID_MIN_U", None)
        assert (ass == wa).all(), f"grid={grid} min_u={min_u}: assignment differs in {(ass != wa).sum()} places"
        assert (dist.view(np.uint32) == wd.view(np.uint32)).all()
    if iters > 1 and kind != "lattice":  # the last iteration's price/assignment_inv writes race in the reference too
        pass


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("cfg", [("uniform", 3, 1024, 0.05, 3000), ("clustered", 2, 2048, 0.05, 3000), ("uniform", 2, 1024, 0.05, 150),
                                 ("uniform", 2, 1024, 0.05, 60), ("uniform", 1, 4096, 0.05, 500), ("dup", 2, 1024, 0.05, 3000)])
def test_emd_solo_mode_training_setting(pkg, oracle, cuda, cluster, cfg):
    """The training setting (loss/loss.py:18-28: eps 0.05, 3000 iterations) spends most iterations with a handful of
    bidders; below 33 of them one CTA finishes the auction alone (solo mode).  Same bits as the oracle and as the
    cluster-wide form, also when the run ends (forced assignment) inside the solo phase."""
    kind, b, n, eps, iters = cfg
    x, y = make_clouds(kind, b, n, n, seed=7 + n + iters)
    wd, wa, state = oracle.emd_forward(x, y, eps, iters, nthreads=8, full_state=True)
    lib = pkg._lib.lib
    res = {}
    for solo, grid in ((1, 1), (0, 1), (1, 0), (0, 0)):
        old, oldg = lib.psd_emd_solo_mode(solo), lib.psd_emd_grid_mode(grid)
        try:
            out = run_cluster(pkg, cuda, x, y, eps, iters, cluster)
        finally:
            lib.psd_emd_solo_mode(old); lib.psd_emd_grid_mode(oldg)
        if grid:
            res[solo] = out
        dist, ass, price, inv = out
        assert (ass == wa).all(), f"solo={solo} grid={grid}: assignment differs in {(ass != wa).sum()} places"
        assert (dist.view(np.uint32) == wd.view(np.uint32)).all()
    if iters == 3000:   # converged long before the last iteration: no forced assignment, the scratch state is deterministic
        for k in (2, 3):
            assert (res[1][k].view(np.uint32) == res[0][k].view(np.uint32)).all()
        assert (res[1][2].view(np.uint32) == state["price"].view(np.uint32)).all()
        assert (res[1][3] == state["assignment_inv"]).all()


def test_emd_module_config3_and_backward(pkg, oracle, cuda):
    """BASELINE.json configs[2]: B=32, n=2048, eps=0.005, iters=50 through emdModule, plus the gradient."""
    x, y = make_clouds("uniform", 32, 2048, 2048, seed=0)
    tx = torch.from_numpy(x).to(cuda).requires_grad_(True)
    ty = torch.from_numpy(y).to(cuda)
    dist, ass = pkg.emdModule()(tx, ty, 0.005, 50)
    wd, wa, st = oracle.emd_forward(x, y, 0.005, 50, nthreads=16, want_stats=True)
    assert (ass.cpu().numpy() == wa).all()
    assert (dist.detach().cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
    print("oracle stats: sum_u", st["sum_u"], "multi_winner", st["multi_winner"], "best==better", st["best_eq_better"])
    g = np.random.default_rng(1).random((32, 2048), dtype=np.float32)
    (dist * torch.from_numpy(g).to(cuda)).sum().backward()
    want = oracle.emd_backward(x, y, g, wa)
    assert (tx.grad.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all()


def test_emd_native_module_signature(pkg, oracle, cuda):
    """emd.forward with the reference's 14 tensors + eps + iters (emd_module.py:43-73) and its return codes."""
    b, n = 2, 1024
    x, y = make_clouds("uniform", b, n, n, seed=3)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=cuda, dtype=dt)
    dist = z(b, n); assignment = z(b, n, dt=torch.int32) - 1; assignment_inv = z(b, n, dt=torch.int32) - 1
    price = z(b, n); bid = z(b, n, dt=torch.int32); bid_increments = z(b, n); max_increments = z(b, n)
    unass_idx = z(b * n, dt=torch.int32); max_idx = z(b * n, dt=torch.int32)
    unass_cnt = z(512, dt=torch.int32); unass_cnt_sum = z(512, dt=torch.int32); cnt_tmp = z(512, dt=torch.int32)
    rc = pkg.emd.forward(tx, ty, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx,
                         unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, 0.005, 50)
    assert rc == 1
    wd, wa = oracle.emd_forward(x, y, 0.005, 50)[:2]
    assert (assignment.cpu().numpy() == wa).all() and (dist.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
    # shape violations return -1 like emd_cuda.cu:236-249
    bad = torch.zeros(2, 1000, 3, device=cuda)
    assert pkg.emd.forward(bad, bad, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments,
                           unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, 0.005, 5) == -1
    with pytest.raises(AssertionError):
        pkg.emdModule()(bad, bad, 0.005, 5)


def test_emd_self_consistency(pkg, cuda):
    """The reference's own check (metric/emd/test.py:24-28): dist == squared distance along the assignment."""
    x, y = make_clouds("uniform", 4, 2048, 2048, seed=8)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    dist, ass = pkg.emdModule()(tx, ty, 0.05, 100)
    sel = torch.gather(ty, 1, ass.long().unsqueeze(-1).expand(-1, -1, 3))
    d = ((tx - sel) ** 2).sum(-1)
    assert torch.allclose(d, dist, rtol=1e-5, atol=1e-7)
    assert int(ass.min()) >= 0 and int(ass.max()) < 2048
