"""GPU parity of the layout seam and the loss / metric glue: transposed [B,3,N] views
read in place through the drop-in API, the backward that needs no zero fill, Loss.get_emd_loss / Metrics.get against the
CPU oracle, and the differentiable cont_proj."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


def _bits_equal(g, w):
    if g.dtype == np.float32:
        return ((g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w))).all()
    return (g == w).all()


@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("shape", [(3, 777, 1029), (24, 1024, 1024), (2, 2500, 700)])
def test_dropin_reads_transposed_views_in_place(pkg, oracle, cuda, variant, shape):
    """train.py:163 passes fake.transpose(2,1) (a view of the generator's [B,3,N] output) and a contiguous ground truth:
    chamfer_3DDist gives the oracle's bits and gradients for every combination of view / contiguous inputs, and the gradient of
    a view comes back in the view's own memory layout (no transpose copy in either direction)."""
    b, n, m = shape
    x, y = make_clouds("uniform", b, n, m, seed=41)
    want = oracle.chamfer_forward(x, y, nthreads=8)
    rng = np.random.default_rng(3)
    g1 = rng.random((b, n), dtype=np.float32)
    g2 = rng.random((b, m), dtype=np.float32)
    w1, w2 = oracle.chamfer_backward(x, y, g1, g2, want[2], want[3])
    old = pkg._lib.lib.psd_chamfer_nn_variant(variant)
    try:
        for view1 in (True, False):
            for view2 in (False, True):
                base1 = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1)) if view1 else x).to(cuda).requires_grad_(True)
                base2 = torch.from_numpy(np.ascontiguousarray(y.transpose(0, 2, 1)) if view2 else y).to(cuda).requires_grad_(True)
                a = base1.transpose(2, 1) if view1 else base1
                c = base2.transpose(2, 1) if view2 else base2
                assert a.is_contiguous() != view1 or n == 1
                d1, d2, i1, i2 = pkg.chamfer_3DDist()(a, c)
                ((d1 * torch.from_numpy(g1).to(cuda)).sum() + (d2 * torch.from_numpy(g2).to(cuda)).sum()).backward()
                torch.cuda.synchronize()
                for got, w in zip((d1, d2, i1, i2), want):
                    assert _bits_equal(got.detach().cpu().numpy(), w), (view1, view2)
                ga = base1.grad.transpose(2, 1) if view1 else base1.grad
                gc = base2.grad.transpose(2, 1) if view2 else base2.grad
                for got, w in ((ga, w1), (gc, w2)):
                    tol = 1e-5 * max(np.abs(w).max(), 1e-30)
                    assert np.abs(got.cpu().numpy() - w).max() <= tol, (view1, view2)
    finally:
        pkg._lib.lib.psd_chamfer_nn_variant(old)


def test_backward_overwrite_ignores_buffer_contents(pkg, oracle, cuda):
    """psd_chamfer_backward_ex(overwrite=1): NaN-filled gradient buffers come back as the oracle's gradients (no zero fill
    needed), also when a thread owns several points (more points than co-resident threads) and for ragged sizes."""
    lib = pkg._lib.lib
    for b, n, m in ((2, 256, 300), (3, 1000, 2000), (40, 4096, 4100), (1, 1, 7)):
        x, y = make_clouds("clustered" if n > 1 else "uniform", b, n, m, seed=21)
        tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
        d1, d2, i1, i2 = pkg.chamfer_3DDist()(tx, ty)
        g1 = torch.rand_like(d1)
        g2 = torch.rand_like(d2)
        a1 = torch.full_like(tx, float("nan"))
        a2 = torch.full_like(ty, float("nan"))
        rc = lib.psd_chamfer_backward_ex(*[ctypes.c_void_p(t.data_ptr()) for t in (tx, ty, a1, a2, g1, g2, i1, i2)], b, n, m, 0, 1, None)
        assert rc == 1, pkg._lib.last_error()
        torch.cuda.synchronize()
        w1, w2 = oracle.chamfer_backward(x, y, g1.cpu().numpy(), g2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy())
        for got, w in ((a1, w1), (a2, w2)):
            tol = 1e-5 * max(np.abs(w).max(), 1e-30)
            assert np.abs(got.cpu().numpy() - w).max() <= tol, (b, n, m)
    # an empty cloud: nothing to pair with -> zeros (overwrite) / untouched (accumulate)
    tx = torch.rand(2, 5, 3, device=cuda)
    a1 = torch.full_like(tx, float("nan"))
    rc = lib.psd_chamfer_backward_ex(ctypes.c_void_p(tx.data_ptr()), None, ctypes.c_void_p(a1.data_ptr()), None, None, None, None, None,
                                     2, 5, 0, 0, 1, None)
    assert rc == 1 and float(a1.abs().sum()) == 0.0
    assert lib.psd_chamfer_backward_ex(None, None, None, None, None, None, None, None, 1, 1, 1, 7, 1, None) == -1   # bad layout


def test_second_backward_through_the_same_graph(pkg, oracle, cuda):
    """The forward launch zero-fills the gradient buffers once; backward(retain_graph=True) twice must not accumulate the
    first pass into the second (fresh zeros the second time), for the module and for the fused loss."""
    x, y = make_clouds("uniform", 2, 300, 400, seed=4)
    want = oracle.chamfer_forward(x, y)
    g1 = np.full((2, 300), 1.0 / 600, np.float32); g2 = np.full((2, 400), 1.0 / 800, np.float32)
    w1, w2 = oracle.chamfer_backward(x, y, g1, g2, want[2], want[3])
    for fused in (False, True):
        tx = torch.from_numpy(x).to(cuda).requires_grad_(True)
        ty = torch.from_numpy(y).to(cuda).requires_grad_(True)
        if fused:
            loss = pkg.Loss().get_chamfer_loss(tx, ty)
        else:
            d1, d2, _, _ = pkg.chamfer_3DDist()(tx, ty)
            loss = d1.mean() + d2.mean()
        (a1, a2) = torch.autograd.grad(loss, (tx, ty), retain_graph=True)
        a1, a2 = a1.clone(), a2.clone()
        (b1, b2) = torch.autograd.grad(loss, (tx, ty))
        for got, w in ((a1, w1), (b1, w1), (a2, w2), (b2, w2)):
            assert np.abs(got.cpu().numpy() - w).max() <= 1e-5 * np.abs(w).max(), fused
    # no gradient wanted: no buffer is allocated or filled
    with torch.no_grad():
        out = pkg.chamfer_3DDist()(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda))
    assert np.array_equal(out[2].cpu().numpy(), want[2])


def test_loss_chamfer_on_transposed_prediction(pkg, oracle, cuda):
    """Loss.get_chamfer_loss(fake.transpose(2,1), points) -- train.py:163 verbatim -- on the fused path without a copy."""
    b, n = 4, 1024
    x, y = make_clouds("uniform", b, n, n, seed=5)
    fake = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).to(cuda).requires_grad_(True)   # [B,3,N]
    points = torch.from_numpy(y).to(cuda)
    loss = pkg.Loss().get_chamfer_loss(fake.transpose(2, 1), points)
    (100.0 * loss).backward()
    torch.cuda.synchronize()
    d1, d2, i1, i2 = oracle.chamfer_forward(x, y, nthreads=8)
    want = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
    assert abs(float(loss) - want) <= 1e-5 * want
    gd1 = np.full((b, n), np.float32(100.0) / np.float32(b * n), np.float32)
    w1, _ = oracle.chamfer_backward(x, y, gd1, gd1.copy(), i1, i2)
    got = fake.grad.transpose(2, 1).cpu().numpy()
    assert fake.grad.shape == (b, 3, n) and np.abs(got - w1).max() <= 1e-5 * np.abs(w1).max()


def test_loss_rejects_wrong_last_dimension(pkg, cuda):
    """a forgotten transpose ([B,3,N] passed as [B,N,3]) must raise like the reference's assert
    (dist_chamfer_3D.py:33-35), not produce a silently wrong loss."""
    bad = torch.rand(2, 3, 64, device=cuda)
    good = torch.rand(2, 64, 3, device=cuda)
    with pytest.raises(AssertionError):
        pkg.Loss().get_chamfer_loss(bad, good)
    with pytest.raises(AssertionError):
        pkg.Loss().get_chamfer_loss(good, torch.rand(2, 64, 4, device=cuda))


@pytest.mark.parametrize("eps,iters,n", [(0.005, 50, 2048), (0.05, 300, 1024)])
def test_loss_emd_value_and_gradient(pkg, oracle, cuda, eps, iters, n):
    """Loss.get_emd_loss (loss/loss.py:18-28): sqrt(dist).mean(1).mean() within 1e-5 of the oracle, the gradient within 1e-5
    of the oracle's emd backward fed with autograd's grad_dist = ((g/B)/n) / (2 sqrt(dist)); xyz2 gets zeros."""
    b = 3
    x, y = make_clouds("uniform", b, n, n, seed=17)
    tx = torch.from_numpy(x).to(cuda).requires_grad_(True)
    ty = torch.from_numpy(y).to(cuda).requires_grad_(True)
    loss = pkg.Loss().get_emd_loss(tx, ty, eps=eps, iters=iters)
    (100.0 * loss).backward()
    torch.cuda.synchronize()
    wd, wa = oracle.emd_forward(x, y, eps, iters, nthreads=4)[:2]
    want = np.sqrt(wd.astype(np.float64)).mean(1).mean()
    assert abs(float(loss) - want) <= 1e-5 * want
    gd = ((np.float32(100.0) / np.float32(b)) / np.float32(n)) / (np.float32(2.0) * np.sqrt(wd))
    wg = oracle.emd_backward(x, y, gd.astype(np.float32), wa)
    got = tx.grad.cpu().numpy()
    assert np.abs(got - wg).max() <= 1e-5 * np.abs(wg).max()
    assert float(ty.grad.abs().sum()) == 0.0
    # the unfused composition of the module gives the same numbers
    tx2 = torch.from_numpy(x).to(cuda).requires_grad_(True)
    dist, _ = pkg.emdModule()(tx2, torch.from_numpy(y).to(cuda), eps, iters)
    l2 = torch.sqrt(dist).mean(1).mean()
    (100.0 * l2).backward()
    assert abs(float(l2) - float(loss)) <= 1e-6 * float(loss)
    assert np.abs(tx2.grad.cpu().numpy() - got).max() <= 1e-5 * np.abs(got).max()


def test_loss_emd_coincident_point_gives_the_references_infinite_factor(pkg, cuda):
    """d sqrt(0) is infinite: a point that coincides with its assigned object makes the reference's loss gradient
    inf * 0 = NaN for that point.  The fused backward reproduces that, and only there."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 1024, 3, generator=g)
    y = torch.rand(1, 1024, 3, generator=g)
    y[0, 5] = x[0, 9]
    tx = x.to(cuda).requires_grad_(True)
    loss = pkg.Loss().get_emd_loss(tx, y.to(cuda), eps=0.005, iters=50)
    loss.backward()
    tx2 = x.to(cuda).requires_grad_(True)
    dist, ass = pkg.emdModule()(tx2, y.to(cuda), 0.005, 50)
    torch.sqrt(dist).mean(1).mean().backward()
    assert int(ass[0, 9]) == 5 and float(dist[0, 9]) == 0.0
    assert torch.equal(torch.isnan(tx.grad), torch.isnan(tx2.grad)) and bool(torch.isnan(tx.grad[0, 9]).all())
    assert int(torch.isnan(tx.grad).any(-1).sum()) == 1


def test_metrics_get_on_the_eval_setting(pkg, oracle, cuda):
    """Metrics.get (utils/metrics.py:29-37, 48-60): [EMD x100 at eps=0.005 / 50 iterations, chamfer x100], batch size 1 like
    testnet.py:92 and a small batch."""
    for b in (1, 3):
        x, y = make_clouds("clustered", b, 2048, 2048, seed=23 + b)
        got = pkg.Metrics.get(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda))
        wd = oracle.emd_forward(x, y, 0.005, 50, nthreads=4)[0]
        d1, d2, _, _ = oracle.chamfer_forward(x, y, nthreads=8)
        want = [np.sqrt(wd.astype(np.float64)).mean(1).mean() * 100, (d1.astype(np.float64).mean() + d2.astype(np.float64).mean()) * 100]
        assert len(got) == 2 and all(isinstance(v, float) for v in got)
        assert abs(got[0] - want[0]) <= 1e-5 * want[0] and abs(got[1] - want[1]) <= 1e-5 * want[1]
        assert abs(pkg.Metrics._get_emd_distance(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)) - want[0]) <= 1e-5 * want[0]
        assert abs(pkg.Metrics._get_chamfer_distance(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)) - want[1]) <= 1e-5 * want[1]


def _ref_cont_proj(pcl, grid_h, grid_w, sigma_sq):
    """The reference's expression (utils/projection.py:18-64) in plain differentiable torch ops."""
    x = (pcl[..., 0:1] + 1) * grid_h / 2
    y = (pcl[..., 1:2] + 1) * grid_w / 2
    xy = torch.cat([x, y], 2)
    gh, gw = torch.meshgrid(torch.arange(grid_h, device=pcl.device, dtype=torch.float32),
                            torch.arange(grid_w, device=pcl.device, dtype=torch.float32), indexing="ij")
    grid = torch.stack([gh, gw], 2)
    diff = xy[:, :, None, None, :] - grid
    val = torch.exp(-(diff ** 2) / (2. * sigma_sq))
    return (val[..., 0] * val[..., 1]).sum(1)


@pytest.mark.parametrize("shape", [(2, 300, 32, 48), (3, 1024, 64, 64), (1, 77, 70, 64)])
def test_cont_proj_is_differentiable_like_the_reference(pkg, cuda, shape):
    """cont_proj must carry gradients.  Value and d/d pcl against autograd through the reference's own
    torch expression: a weighted sum of the silhouette (a stand-in for the BCE projection loss of finetune.py:158-162)."""
    b, n, h, w = shape
    g = torch.Generator().manual_seed(9)
    pcl = (torch.rand(b, n, 3, generator=g) * 1.9 - 0.95).to(cuda)
    wgt = torch.rand(b, h, w, generator=g).to(cuda)
    p1 = pcl.clone().requires_grad_(True)
    out = pkg.projection.cont_proj(p1, h, w, cuda, 0.5)
    assert out.requires_grad
    (out * wgt).sum().backward()
    p2 = pcl.clone().requires_grad_(True)
    ref = _ref_cont_proj(p2, h, w, 0.5)
    (ref * wgt).sum().backward()
    assert torch.allclose(out, ref, rtol=2e-5, atol=1e-6)
    scale = float(p2.grad.abs().max())
    assert float((p1.grad - p2.grad).abs().max()) <= 2e-5 * scale
    assert float(p1.grad[..., 2].abs().max()) == 0.0
    # a CPU input keeps its gradient path through the device copy
    p3 = pcl.cpu().clone().requires_grad_(True)
    pkg.projection.cont_proj(p3, h, w, "cpu", 0.5).sum().backward()
    assert p3.grad is not None and float(p3.grad.abs().max()) > 0


def test_host_step_with_the_prediction_on_the_device(pkg, oracle, cuda):
    """psd_chamfer_loss_step_pred_dev: the generator's output ([B,3,N], device) against a ground truth on the host
    (train.py:160-163): loss and d loss / d pred, plain call and cached-graph replay."""
    lib = pkg._lib.lib
    b, n, m = 6, 1024, 1024
    x, y = make_clouds("uniform", b, n, m, seed=61)
    pred = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).to(cuda)          # [B,3,N]
    gt = torch.from_numpy(y).pin_memory()
    loss = torch.zeros(1).pin_memory()
    grad = torch.full((b, 3, n), float("nan"), device=cuda)
    d1, d2, i1, i2 = oracle.chamfer_forward(x, y, nthreads=8)
    want = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
    gd1 = np.full((b, n), np.float32(1.0) / np.float32(b * n), np.float32)
    gd2 = np.full((b, m), np.float32(1.0) / np.float32(b * m), np.float32)
    w1, _ = oracle.chamfer_backward(x, y, gd1, gd2, i1, i2)
    st = torch.cuda.Stream()
    for rep in range(3):   # the third call replays the cached CUDA graph
        grad.fill_(float("nan"))
        torch.cuda.synchronize()
        rc = lib.psd_chamfer_loss_step_pred_dev(ctypes.c_void_p(pred.data_ptr()), 1, ctypes.c_void_p(gt.data_ptr()), b, n, m,
                                                ctypes.c_void_p(loss.data_ptr()), ctypes.c_void_p(grad.data_ptr()), 1, 1,
                                                ctypes.c_void_p(st.cuda_stream))
        assert rc == 1, pkg._lib.last_error()
        assert abs(float(loss[0]) - want) <= 1e-5 * want
        got = grad.transpose(1, 2).cpu().numpy()
        assert np.abs(got - w1).max() <= 1e-5 * np.abs(w1).max(), rep


def test_proj_min_dist_terms_refuse_to_drop_a_gradient(pkg, cuda):
    """the min-distance terms have no backward kernel.  With inputs that require grad the values are still right and
    usable for logging, but differentiating THROUGH them raises instead of silently contributing nothing."""
    g = torch.Generator().manual_seed(2)
    pred = torch.rand(2, 16, 16, generator=g).to(cuda).requires_grad_(True)
    gt = (torch.rand(2, 16, 16, generator=g) > 0.5).float().to(cuda)
    dm = torch.from_numpy(pkg.proj_loss.grid_dist(16, 16)).float()
    loss, fwd, bwd = pkg.proj_loss.get_loss_proj(pred, gt, cuda, "bce_prob", 1.0, True, dm)
    want_f, want_b = pkg.proj_loss.min_dist_terms(pred.detach(), gt, dm)
    assert torch.equal(fwd.detach(), want_f) and torch.equal(bwd.detach(), want_b) and fwd.requires_grad
    loss.backward()                                    # the BCE term is plain torch and differentiable
    assert pred.grad is not None
    with pytest.raises(NotImplementedError):
        (1e-4 * fwd.mean()).backward()
