"""EMD beyond the shared-memory resident size (VERDICT r1 missing 3): the reference accepts any n % 1024 == 0
(metric/emd/emd_cuda.cu:125-133,236-249); clouds of more than 8192 points run the global-workspace form of the auction
kernel.  Bit-exact against the CPU oracle; the form is also forced at small n for every cluster size."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


def _run_cluster(pkg, cuda, x, y, eps, iters, cluster):
    b, n, _ = x.shape
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    dist = torch.zeros(b, n, device=cuda)
    ass = torch.full((b, n), -1, device=cuda, dtype=torch.int32)
    price = torch.zeros(b, n, device=cuda)
    inv = torch.full((b, n), -1, device=cuda, dtype=torch.int32)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = pkg._lib.lib.psd_emd_forward_cluster(vp(tx), vp(ty), b, n, vp(dist), vp(ass), vp(price), vp(inv), eps, iters, cluster, None)
    assert rc == 1, pkg._lib.last_error()
    torch.cuda.synchronize()
    return dist.cpu().numpy(), ass.cpu().numpy(), price.cpu().numpy(), inv.cpu().numpy()


@pytest.mark.parametrize("cluster", [-1, -2, -4, -8])
@pytest.mark.parametrize("kind,n,eps,iters", [("uniform", 2048, 0.005, 50), ("clustered", 1024, 0.05, 200), ("dup", 1024, 0.005, 30)])
def test_global_workspace_form_matches_oracle(pkg, oracle, cuda, cluster, kind, n, eps, iters):
    x, y = make_clouds(kind, 3, n, n, seed=99)
    want = oracle.emd_forward(x, y, eps, iters, nthreads=4)
    got = _run_cluster(pkg, cuda, x, y, eps, iters, cluster)
    assert np.array_equal(got[1], want[1]), "assignment"
    assert np.array_equal(got[0].view(np.uint32), want[0].view(np.uint32)), "dist"


@pytest.mark.parametrize("n,b,iters", [(16384, 2, 6), (9216, 1, 12)])
def test_large_clouds_through_the_module(pkg, oracle, cuda, n, b, iters):
    """n = 16384 (and 9216 = 9 x 1024, not a power of two) through emdModule: returns, and dist / assignment are bit-exact
    (few iterations keep the CPU oracle at seconds: the work is sum_t u_t * n)."""
    x, y = make_clouds("uniform", b, n, n, seed=7)
    dist, ass = pkg.emdModule()(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda), 0.005, iters)
    torch.cuda.synchronize()
    wd, wa = oracle.emd_forward(x, y, 0.005, iters, nthreads=8)[:2]
    assert np.array_equal(ass.cpu().numpy(), wa)
    assert np.array_equal(dist.cpu().numpy().view(np.uint32), wd.view(np.uint32))
    # every point is assigned after the last (forcing) iteration and dist is the squared distance along the assignment
    a = ass.long()
    sel = torch.gather(torch.from_numpy(y).to(cuda), 1, a.unsqueeze(-1).expand(-1, -1, 3))
    assert int(a.min()) >= 0 and torch.allclose(((torch.from_numpy(x).to(cuda) - sel) ** 2).sum(-1), dist, rtol=1e-5, atol=1e-9)


def test_large_cloud_raw_signature_and_gradient(pkg, oracle, cuda):
    """The reference's 16-argument emd.forward at n = 16384 (caller-initialised state as emd_module.py:43-54) returns 1, and
    the fused loss backward works on the result."""
    b, n = 1, 16384
    x, y = make_clouds("uniform", b, n, n, seed=8)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=cuda, dtype=dt)
    dist, ass, price, inv = z(b, n), z(b, n, dt=torch.int32) - 1, z(b, n), z(b, n, dt=torch.int32) - 1
    rest = [z(b, n, dt=torch.int32), z(b, n), z(b, n), z(b * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
            z(512, dt=torch.int32), z(b * n, dt=torch.int32)]
    assert pkg.emd.forward(tx, ty, dist, ass, price, inv, *rest, 0.005, 4) == 1
    torch.cuda.synchronize()
    want = oracle.emd_forward(x, y, 0.005, 4, nthreads=8)
    assert np.array_equal(ass.cpu().numpy(), want[1]) and np.array_equal(dist.cpu().numpy().view(np.uint32), want[0].view(np.uint32))
    txg = tx.clone().requires_grad_(True)
    loss = pkg.Loss().get_emd_loss(txg, ty, eps=0.005, iters=4)
    loss.backward()
    assert abs(float(loss) - np.sqrt(want[0].astype(np.float64)).mean()) <= 1e-5 * float(loss)
    assert bool(torch.isfinite(txg.grad).all())
