"""GPU parity against the REFERENCE ITSELF: the reference's chamfer3D / emd CUDA extensions compiled unmodified
into oracle/_ref (oracle/build_ref.py) run next to this library on identical inputs.  Skipped when the
prebuilt modules are absent (they are built in the build container and shipped with the snapshot)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    for d in ("ref_chamfer_3D", "ref_emd"):
        p = os.path.join(ROOT, "oracle", "_ref", d)
        if not os.path.isdir(p) or not any(f.endswith(".so") for f in os.listdir(p)):
            pytest.skip("oracle/_ref not built (run python oracle/build_ref.py where /root/reference exists)")
        sys.path.insert(0, p)
    import ref_chamfer_3D
    import ref_emd
    return ref_chamfer_3D, ref_emd


@pytest.mark.parametrize("kind,shape", [("uniform", (32, 2048, 2048)), ("uniform", (32, 1000, 2000)), ("lattice", (4, 700, 900)),
                                        ("dup", (3, 1024, 1536)), ("clustered", (8, 2048, 1024))])
def test_chamfer_matches_reference_extension(pkg, oracle, cuda, ref, kind, shape):
    b, n, m = shape  # (32,1000,2000) is the reference's own smoke shape, metric/chamfer3D/test.py:4-5
    x, y = make_clouds(kind, b, n, m, seed=42)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    rd1 = torch.zeros(b, n, device=cuda); rd2 = torch.zeros(b, m, device=cuda)
    ri1 = torch.zeros(b, n, device=cuda, dtype=torch.int32); ri2 = torch.zeros(b, m, device=cuda, dtype=torch.int32)
    assert ref[0].forward(tx, ty, rd1, rd2, ri1, ri2) == 1
    d1, d2, i1, i2 = pkg.chamfer_3DDist()(tx, ty)
    torch.cuda.synchronize()
    assert torch.equal(i1, ri1) and torch.equal(i2, ri2), "idx differs from the reference extension"
    assert torch.equal(d1, rd1) and torch.equal(d2, rd2), "dist differs from the reference extension"
    w = oracle.chamfer_forward(x, y, nthreads=16)
    assert np.array_equal(w[2], ri1.cpu().numpy()) and np.array_equal(w[0], rd1.cpu().numpy()), "oracle != reference"
    # gradients: both sides accumulate with float atomics -> 1e-5 relative
    g1 = torch.rand(b, n, device=cuda); g2 = torch.rand(b, m, device=cuda)
    ra = torch.zeros_like(tx); rb = torch.zeros_like(ty)
    assert ref[0].backward(tx, ty, ra, rb, g1, g2, ri1, ri2) == 1
    oa = torch.zeros_like(tx); ob = torch.zeros_like(ty)
    assert pkg.chamfer_3D.backward(tx, ty, oa, ob, g1, g2, i1, i2) == 1
    torch.cuda.synchronize()
    for got, want in ((oa, ra), (ob, rb)):
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize("cfg", [(32, 2048, 0.005, 50), (8, 1024, 0.05, 300), (20, 2048, 0.05, 50)])
def test_emd_matches_reference_extension(pkg, oracle, cuda, ref, cfg):
    b, n, eps, iters = cfg  # (20, 2048, 0.05, 50) is the reference's own smoke call, metric/emd/test.py:7-12
    x, y = make_clouds("uniform", b, n, n, seed=7)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=cuda, dtype=dt)
    rdist = z(b, n); rass = z(b, n, dt=torch.int32) - 1; rinv = z(b, n, dt=torch.int32) - 1; rprice = z(b, n)
    args = [z(b, n, dt=torch.int32), z(b, n), z(b, n), z(b * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
            z(512, dt=torch.int32), z(b * n, dt=torch.int32)]
    assert ref[1].forward(tx, ty, rdist, rass, rprice, rinv, *args, eps, iters) == 1
    dist, ass = pkg.emdModule()(tx, ty, eps, iters)
    torch.cuda.synchronize()
    wd, wa, st = oracle.emd_forward(x, y, eps, iters, nthreads=16, want_stats=True)
    assert np.array_equal(ass.cpu().numpy(), wa) and np.array_equal(dist.cpu().numpy(), wd)
    # the reference's GetMax is a last-writer-wins race: require identity on clouds without a multi-winner
    # event, and the loss within 1e-5 relative overall (SURVEY.md section 7 parity protocol)
    same = (rass == ass).all(1).cpu().numpy()
    print(f"clouds identical to the reference extension: {int(same.sum())}/{b}; oracle multi-winner events: {st['multi_winner']}")
    assert int((~same).sum()) <= max(st["multi_winner"], 0) + 2
    # clouds without a race in the reference: bit-identical distances
    sm = torch.from_numpy(same).to(cuda)
    assert torch.equal(rdist[sm], dist[sm])
    # clouds where the reference's race picked another (equally valid) winner: the auction takes a different
    # path from there on, so only the per-cloud loss is comparable -- both are eps-approximations of the same EMD
    if (~sm).any():
        rl = torch.sqrt(rdist[~sm]).mean(1); ol = torch.sqrt(dist[~sm]).mean(1)
        assert float(((rl - ol).abs() / rl).max()) <= 2e-2


@pytest.mark.parametrize("kind,b", [("uniform", 16), ("lattice", 6)])
def test_emd_n4096_tie_order_versus_reference_extension(pkg, oracle, cuda, ref, kind, b, record_property):
    """n = 4096 = two 2048-object tiles in the reference's Bid (emd_cuda.cu:125-158): a bidder's equal best values are
    resolved by its (thread, tile) scan order there, by the LOWEST object index here and in the oracle (the north-star rule,
    DESIGN section 2).  This test MEASURES how often that matters: it counts the clouds whose assignment differs from the
    reference extension and, with the oracle's diagnostics (bids whose best equals their second best, multi-winner events),
    attributes them.  Uniform clouds: exact value ties essentially never occur, the count is bounded by the GetMax race;
    lattice clouds (coordinates on a 1/8 grid: massive exact ties) differ wholesale -- only the loss is comparable there."""
    n, eps, iters = 4096, 0.005, 30
    x, y = make_clouds(kind, b, n, n, seed=11)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=cuda, dtype=dt)
    rdist = z(b, n); rass = z(b, n, dt=torch.int32) - 1; rinv = z(b, n, dt=torch.int32) - 1; rprice = z(b, n)
    args = [z(b, n, dt=torch.int32), z(b, n), z(b, n), z(b * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
            z(512, dt=torch.int32), z(b * n, dt=torch.int32)]
    assert ref[1].forward(tx, ty, rdist, rass, rprice, rinv, *args, eps, iters) == 1
    dist, ass = pkg.emdModule()(tx, ty, eps, iters)
    torch.cuda.synchronize()
    wd, wa, st = oracle.emd_forward(x, y, eps, iters, nthreads=16, want_stats=True)
    assert np.array_equal(ass.cpu().numpy(), wa) and np.array_equal(dist.cpu().numpy(), wd)      # ours == oracle, always
    same = (rass == ass).all(1).cpu().numpy()
    differing = int((~same).sum())
    rl = torch.sqrt(rdist).mean(1); ol = torch.sqrt(dist).mean(1)
    rel = float(((rl - ol).abs() / rl).max())
    msg = (f"EMD n=4096 {kind}: {differing}/{b} clouds differ from the reference extension; oracle: {st['best_eq_better']} bids with "
           f"best == second best (exact value ties), {st['multi_winner']} multi-winner events; max relative loss difference {rel:.2e}")
    print(msg)
    record_property("tie_order_report", msg)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "emd_n4096_tie_order.txt"), "a") as f:
            f.write(msg + "\n")
    except OSError:
        pass
    if kind == "uniform":
        # no exact value ties -> the tie order cannot matter; what differs is bounded by the reference's GetMax race
        assert differing <= st["multi_winner"] + st["best_eq_better"] + 2
        sm = torch.from_numpy(same).to(cuda)
        assert torch.equal(rdist[sm], dist[sm])
    assert rel <= 5e-2      # both are eps-approximations of the same EMD
