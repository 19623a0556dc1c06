"""Out-of-bounds evidence without compute-sanitizer (the tool is closed on this GPU pool, profiles/r2_compute_sanitizer_closed.txt):
every tensor an entry point of the C ABI touches is carved out of a larger allocation with sentinel-filled guard zones on both
sides; after the call the guards must be intact and the inputs unchanged.  Covers every kernel family at ragged sizes (the
sizes where an index bug shows), both NN kernels, multi-tile mode, both layouts, every EMD decomposition."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu
G = 2048            # guard elements on each side
SENT = 0x7FC0DEAD   # a NaN pattern as float32, an unmistakable value as int32


class Arena:
    def __init__(self, dev):
        self.dev, self.items = dev, []

    def _new(self, numel, dtype):
        raw = torch.full((numel + 2 * G,), SENT, dtype=torch.int32, device=self.dev)
        self.items.append((raw, numel))
        return raw[G:G + numel].view(dtype)

    def tensor(self, src):
        """guarded device copy of a numpy array / tensor (4-byte dtypes)"""
        t = torch.as_tensor(src)
        out = self._new(t.numel(), t.dtype)
        out.copy_(t.reshape(-1).to(self.dev))
        return out.view(t.shape)

    def empty(self, shape, dtype=torch.float32, fill=None):
        out = self._new(int(np.prod(shape)), dtype).view(shape)
        if fill is not None:
            out.fill_(fill)
        return out

    def check(self, what):
        torch.cuda.synchronize()
        for raw, numel in self.items:
            assert bool((raw[:G] == SENT).all()) and bool((raw[G + numel:] == SENT).all()), f"{what}: a guard zone was written"


vp = lambda t: ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("variant", [1, 3])
@pytest.mark.parametrize("shape", [(3, 257, 130), (2, 1, 700), (1, 2049, 4099), (5, 640, 640)])
@pytest.mark.parametrize("layout", [0, 1, 2, 3])
def test_chamfer_entry_points_stay_in_bounds(pkg, cuda, variant, shape, layout):
    lib = pkg._lib.lib
    b, n, m = shape
    x, y = make_clouds("uniform", b, n, m, seed=3)
    A = Arena(cuda)
    xin = np.ascontiguousarray(x.transpose(0, 2, 1)) if layout & 1 else x
    yin = np.ascontiguousarray(y.transpose(0, 2, 1)) if layout & 2 else y
    tx, ty = A.tensor(xin), A.tensor(yin)
    d1, d2 = A.empty((b, n)), A.empty((b, m))
    i1, i2 = A.empty((b, n), torch.int32), A.empty((b, m), torch.int32)
    sums, cnt = A.empty((b, 2), fill=0.0), A.empty((b, 2), torch.int32, fill=0)
    grads = A.empty((3 * b * (n + m),), fill=float("nan"))
    loss = A.empty((1,))
    old = lib.psd_chamfer_nn_variant(variant)
    try:
        assert lib.psd_chamfer_forward_ex(vp(tx), vp(ty), b, n, m, layout, vp(d1), vp(d2), vp(i1), vp(i2), vp(sums), 1e-4, vp(cnt), 0, -1, None) == 1
        A.check("forward_ex")
        q0, qc = max(n, m) // 3, max(n, m) // 2
        assert lib.psd_chamfer_forward_ex(vp(tx), vp(ty), b, n, m, layout, vp(d1), vp(d2), vp(i1), vp(i2), vp(sums), 1e-4, vp(cnt), q0, qc, None) == 1
        A.check("forward_ex, query slice")
        assert lib.psd_chamfer_forward_zero(vp(tx), vp(ty), b, n, m, layout, vp(d1), vp(d2), vp(i1), vp(i2), None, 0.0, None, vp(grads), grads.numel(), None) == 1
        A.check("forward_zero")
        assert float(grads.abs().sum()) == 0.0
        g1, g2 = grads[: 3 * b * n], grads[3 * b * n:]
        gd1, gd2 = A.tensor(np.random.default_rng(1).random((b, n), dtype=np.float32)), A.tensor(np.random.default_rng(2).random((b, m), dtype=np.float32))
        assert lib.psd_chamfer_backward_ex(vp(tx), vp(ty), vp(g1), vp(g2), vp(gd1), vp(gd2), vp(i1), vp(i2), b, n, m, layout, 0, None) == 1
        A.check("backward_ex accumulate")
        keep = grads.clone()
        grads.fill_(float("nan"))
        assert lib.psd_chamfer_backward_ex(vp(tx), vp(ty), vp(g1), vp(g2), vp(gd1), vp(gd2), vp(i1), vp(i2), b, n, m, layout, 1, None) == 1
        A.check("backward_ex overwrite")
        assert torch.allclose(grads, keep, rtol=1e-5, atol=1e-7)
        sums.zero_()
        assert lib.psd_chamfer_mean_loss_forward_zero(vp(tx), vp(ty), b, n, m, layout, vp(d1), vp(d2), vp(i1), vp(i2), vp(sums), vp(loss), vp(grads), grads.numel(), None) == 1
        assert lib.psd_chamfer_mean_loss_backward_ex(vp(tx), vp(ty), vp(g1), vp(g2), None, vp(i1), vp(i2), b, n, m, layout, 0, None) == 1
        A.check("mean loss forward / backward")
    finally:
        lib.psd_chamfer_nn_variant(old)
    assert np.array_equal(tx.cpu().numpy(), xin) and np.array_equal(ty.cpu().numpy(), yin), "an input was modified"
    assert int(i1.min()) >= 0 and int(i1.max()) < m and int(i2.min()) >= 0 and int(i2.max()) < n


@pytest.mark.parametrize("cluster", [0, 1, 2, 4, 8, -1, -8])
def test_emd_entry_points_stay_in_bounds(pkg, cuda, cluster):
    lib = pkg._lib.lib
    b, n = 3, 1024
    x, y = make_clouds("clustered", b, n, n, seed=4)
    A = Arena(cuda)
    tx, ty = A.tensor(x), A.tensor(y)
    dist, ass = A.empty((b, n), fill=0.0), A.empty((b, n), torch.int32, fill=-1)
    price, inv = A.empty((b, n), fill=0.0), A.empty((b, n), torch.int32, fill=-1)
    if cluster == 0:
        bid, binc, minc = A.empty((b, n), torch.int32, fill=0), A.empty((b, n), fill=0.0), A.empty((b, n), fill=0.0)
        assert lib.psd_emd_forward(vp(tx), vp(ty), b, n, n, vp(dist), vp(ass), vp(price), vp(inv), vp(bid), vp(binc), vp(minc),
                                   None, None, None, None, None, 0.05, 120, None) == 1
    else:
        assert lib.psd_emd_forward_cluster(vp(tx), vp(ty), b, n, vp(dist), vp(ass), vp(price), vp(inv), 0.05, 120, cluster, None) == 1
    A.check(f"emd forward, cluster {cluster}")
    assert int(ass.min()) >= 0 and int(ass.max()) < n
    sums, loss, g = A.empty((b,), fill=0.0), A.empty((1,)), A.empty((b, n, 3), fill=float("nan"))
    assert lib.psd_emd_mean_loss_forward(vp(tx), vp(ty), b, n, vp(dist), vp(ass), 0.05, 60, vp(sums), vp(loss), None) == 1
    assert lib.psd_emd_mean_loss_backward(vp(tx), vp(ty), vp(g), vp(dist), vp(ass), None, b, n, None) == 1
    gd = A.tensor(np.random.default_rng(5).random((b, n), dtype=np.float32))
    assert lib.psd_emd_backward_ex(vp(tx), vp(ty), vp(g), vp(gd), vp(ass), b, n, 1, None) == 1
    assert lib.psd_emd_backward(vp(tx), vp(ty), vp(g), vp(gd), vp(ass), b, n, None) == 1
    A.check("emd loss / backward")
    assert np.array_equal(tx.cpu().numpy(), x) and np.array_equal(ty.cpu().numpy(), y)


def test_neighbour_ops_stay_in_bounds(pkg, cuda):
    """FPS, the projection splat with its backward, the projection min-distance kernels, batched ICP and the fp64 NN."""
    lib = pkg._lib.lib
    A = Arena(cuda)
    x, y = make_clouds("uniform", 3, 333, 333, seed=6)
    tx = A.tensor(x)
    cent = torch.full((3 * 40 + 2 * G,), -7, dtype=torch.int64, device=cuda)
    assert lib.psd_farthest_point_sample(vp(tx), 3, 333, 40, 0, ctypes.c_void_p(cent[G:].data_ptr()), None) == 1
    torch.cuda.synchronize()
    assert bool((cent[:G] == -7).all()) and bool((cent[G + 120:] == -7).all())
    p = A.tensor(x * 1.8 - 0.9)
    for h, w in ((33, 47), (64, 64)):
        img, gout, gp = A.empty((3, h, w)), A.tensor(np.random.default_rng(7).random((3, h, w), dtype=np.float32)), A.empty((3, 333, 3))
        assert lib.psd_cont_proj(vp(p), 3, 333, h, w, 0.5, vp(img), None) == 1
        assert lib.psd_cont_proj_backward(vp(p), vp(gout), 3, 333, h, w, 0.5, vp(gp), None) == 1
        table = A.tensor(np.sqrt(np.arange(h)[:, None] ** 2.0 + np.arange(w)[None, :] ** 2.0).astype(np.float32) + 1)
        o1, o2 = A.empty((3, h, w)), A.empty((3, h, w))
        for mode in (0, 1):
            assert lib.psd_proj_min_dist(vp(img), vp(gout), vp(table), 3, h, w, mode, vp(o1), vp(o2), None) == 1
        A.check(f"projection {h}x{w}")
    ty = A.tensor(y)
    T = torch.full((3 * 16 + 2 * G,), -3.0, dtype=torch.float64, device=cuda)
    dd = torch.full((3 * 333 + 2 * G,), -3.0, dtype=torch.float64, device=cuda)
    it = A.empty((3,), torch.int32)
    assert lib.psd_icp_batch(vp(tx), vp(ty), 0, 3, 333, None, 7, 1e-9, ctypes.c_void_p(T[G:].data_ptr()), ctypes.c_void_p(dd[G:].data_ptr()), vp(it), None) == 1
    idx = A.empty((3, 333), torch.int32)
    assert lib.psd_nn_f64(vp(tx), vp(ty), 0, 3, 333, 333, ctypes.c_void_p(dd[G:].data_ptr()), vp(idx), None) == 1
    A.check("icp / nn_f64")
    assert bool((T[:G] == -3.0).all()) and bool((T[G + 48:] == -3.0).all()) and bool((dd[:G] == -3.0).all()) and bool((dd[G + 999:] == -3.0).all())
    assert np.array_equal(tx.cpu().numpy(), x)
