"""GPU parity: CUDA chamfer path (through the C ABI) vs the CPU oracle, bit-exact dist AND idx."""
import numpy as np
import pytest
import torch

from conftest import make_clouds

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 3], ids=["auto", "ffma", "tensor_core"], autouse=True)
def nn_variant(request, pkg):
    """Every test of this file runs against every NN forward kernel (psd_chamfer_nn_variant): the automatic choice, the
    FFMA kernel and the tcgen05 tensor-core kernel (clouds of more than 2048 points: its multi-tile mode)."""
    old = pkg._lib.lib.psd_chamfer_nn_variant(request.param)
    yield request.param
    pkg._lib.lib.psd_chamfer_nn_variant(old)


def run_forward(pkg, dev, x, y):
    tx, ty = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    d1, d2, i1, i2 = pkg.chamfer_3DDist()(tx, ty)
    torch.cuda.synchronize()
    return d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()


def assert_bit_equal(got, want, what):
    for g, w, name in zip(got, want, ("dist1", "dist2", "idx1", "idx2")):
        if g.dtype == np.float32:
            same = (g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w))
        else:
            same = g == w
        assert same.all(), f"{what}: {name} differs at {np.argwhere(~same)[:5].tolist()} ({(~same).sum()} elements)"


@pytest.mark.parametrize("kind", ["uniform", "clustered", "lattice", "dup", "offset"])
@pytest.mark.parametrize("shape", [(2, 1024, 1024), (3, 1000, 2000), (1, 1, 1), (2, 7, 513), (2, 515, 3), (1, 129, 1025), (4, 2048, 2048),
                                   (1, 2500, 4100), (2, 4097, 300), (1, 2049, 6145)])   # > 2048 targets: multi-tile mode of the TC kernel
def test_forward_bit_exact(pkg, oracle, cuda, kind, shape):
    b, n, m = shape
    x, y = make_clouds(kind, b, n, m, seed=1234 + n + m)
    got = run_forward(pkg, cuda, x, y)
    want = oracle.chamfer_forward(x, y, nthreads=8)
    assert_bit_equal(got, want, f"{kind} {shape}")


def test_forward_full_size_config2(pkg, oracle, cuda):
    """BASELINE.json configs[1]: B=32, N=M=2048, against the oracle on all 2.7e8 pairs."""
    x, y = make_clouds("uniform", 32, 2048, 2048, seed=0)
    got = run_forward(pkg, cuda, x, y)
    want = oracle.chamfer_forward(x, y, nthreads=16)
    assert_bit_equal(got, want, "config2")
    fb = np.zeros(2, np.int64)
    import ctypes
    pkg._lib.lib.psd_chamfer_stats(fb.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 0)
    print("fallback queries so far:", fb[1])


def test_forward_identical_clouds_and_all_duplicates(pkg, oracle, cuda):
    x, _ = make_clouds("uniform", 2, 600, 600, seed=5)
    got = run_forward(pkg, cuda, x, x.copy())
    want = oracle.chamfer_forward(x, x.copy())
    assert_bit_equal(got, want, "identical clouds")
    assert (got[0] == 0).all() and (got[2] == np.arange(600)[None]).all()
    y = np.repeat(x[:, :1], 700, axis=1).copy()  # every target is the same point: idx must be 0 everywhere
    got = run_forward(pkg, cuda, x, y)
    want = oracle.chamfer_forward(x, y)
    assert_bit_equal(got, want, "all-duplicate targets")
    assert (got[2] == 0).all()


def test_forward_nan_inf_semantics(pkg, oracle, cuda):
    """Non-finite inputs take the exact path, which follows the reference's 512-tile NaN behaviour."""
    x, y = make_clouds("uniform", 2, 300, 1200, seed=9)
    y[0, 0, 1] = np.nan      # poisons tile 0 of cloud 0 in direction 1
    y[1, 700, 0] = np.nan    # a NaN inside tile 1 of cloud 1: ignored
    y[1, 512, 2] = np.nan    # first element of tile 1 of cloud 1: the whole tile is lost
    x[1, 5, 0] = np.inf
    got = run_forward(pkg, cuda, x, y)
    want = oracle.chamfer_forward(x, y)
    assert_bit_equal(got, want, "nan/inf")


def test_forward_nan_semantics_across_target_tiles(pkg, oracle, cuda):
    """> 2048 targets (multi-tile mode of the tensor-core kernel): NaN at target 0 poisons, NaNs at 512-tile heads in later
    2048-tiles drop their 512-tile, a NaN elsewhere is ignored."""
    x, y = make_clouds("uniform", 2, 400, 5000, seed=19)
    y[0, 0, 2] = np.nan        # target 0 of cloud 0: every query of cloud 0 keeps NaN / index 0
    y[1, 2048, 0] = np.nan     # head of 512-tile 4 (first element of the second 2048-tile) of cloud 1
    y[1, 4608, 1] = np.nan     # head of 512-tile 9
    y[1, 3000, 1] = np.nan     # inside a tile: ignored
    x[1, 7, 1] = np.nan        # a NaN query
    got = run_forward(pkg, cuda, x, y)
    want = oracle.chamfer_forward(x, y)
    assert_bit_equal(got, want, "nan across tiles")


def test_forward_huge_and_tiny_magnitudes(pkg, oracle, cuda):
    x, y = make_clouds("uniform", 2, 256, 700, seed=11)
    for scale in (1e-20, 1e-3, 1e4, 1e12, 3e19):
        xs, ys = (x * np.float32(scale)).astype(np.float32), (y * np.float32(scale)).astype(np.float32)
        got = run_forward(pkg, cuda, xs, ys)
        want = oracle.chamfer_forward(xs, ys)
        assert_bit_equal(got, want, f"scale {scale}")


def test_backward_matches_oracle(pkg, oracle, cuda):
    for kind, (b, n, m) in (("uniform", (4, 1024, 1024)), ("clustered", (3, 1000, 2000)), ("lattice", (2, 333, 77)), ("dup", (2, 512, 640))):
        x, y = make_clouds(kind, b, n, m, seed=77)
        tx = torch.from_numpy(x).to(cuda).requires_grad_(True)
        ty = torch.from_numpy(y).to(cuda).requires_grad_(True)
        d1, d2, i1, i2 = pkg.chamfer_3DDist()(tx, ty)
        rng = np.random.default_rng(3)
        g1 = rng.random((b, n), dtype=np.float32)
        g2 = rng.random((b, m), dtype=np.float32)
        (d1 * torch.from_numpy(g1).to(cuda)).sum().add((d2 * torch.from_numpy(g2).to(cuda)).sum()).backward()
        want1, want2 = oracle.chamfer_backward(x, y, g1, g2, i1.cpu().numpy(), i2.cpu().numpy())
        # fp32 atomics accumulate in arbitrary order: 1e-5 relative (BASELINE.json north_star), scaled per tensor
        for got, want in ((tx.grad.cpu().numpy(), want1), (ty.grad.cpu().numpy(), want2)):
            tol = 1e-5 * max(np.abs(want).max(), 1e-30)
            assert np.abs(got - want).max() <= tol, (kind, np.abs(got - want).max(), tol)


def test_fused_mean_loss_matches_unfused_and_oracle(pkg, oracle, cuda):
    """Loss.get_chamfer_loss (loss/loss.py:30-37) on the fused path: loss within 1e-5 relative of mean(d1)+mean(d2)
    from the oracle, gradients within 1e-5 relative of the oracle's backward fed with autograd's constant gradients
    1/(B*N), 1/(B*M) scaled by an upstream factor."""
    b, n, m = 3, 700, 1100
    x, y = make_clouds("uniform", b, n, m, seed=77)
    tx = torch.from_numpy(x).to(cuda).requires_grad_(True)
    ty = torch.from_numpy(y).to(cuda).requires_grad_(True)
    loss = pkg.Loss().get_chamfer_loss(tx, ty)
    (3.0 * loss).backward()
    torch.cuda.synchronize()
    d1, d2, i1, i2 = oracle.chamfer_forward(x, y, nthreads=8)
    want = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
    assert abs(float(loss) - want) <= 1e-5 * abs(want)
    gd1 = np.full((b, n), np.float32(3.0) / np.float32(b * n), np.float32)
    gd2 = np.full((b, m), np.float32(3.0) / np.float32(b * m), np.float32)
    g1, g2 = oracle.chamfer_backward(x, y, gd1, gd2, i1, i2)
    for got, ref in ((tx.grad.cpu().numpy(), g1), (ty.grad.cpu().numpy(), g2)):
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 1e-5 * scale
    # and the unfused public path gives the same loss
    o1, o2, _, _ = pkg.chamfer_3DDist()(tx.detach(), ty.detach())
    unfused = float(torch.mean(o1) + torch.mean(o2))
    assert abs(float(loss) - unfused) <= 1e-5 * abs(unfused)


def test_host_buffer_loss_step(pkg, oracle, cuda):
    """psd_chamfer_loss_step_host: loss and gradients of one training step with host buffers, against the oracle."""
    b, n, m = 2, 900, 640
    x, y = make_clouds("clustered", b, n, m, seed=5)
    hx, hy = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    loss, g1, g2 = pkg.chamfer_loss_step_host(hx, hy, want_grads=True)
    d1, d2, i1, i2 = oracle.chamfer_forward(x, y, nthreads=8)
    want = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
    assert abs(loss - want) <= 1e-5 * abs(want)
    gd1 = np.full((b, n), np.float32(1.0) / np.float32(b * n), np.float32)
    gd2 = np.full((b, m), np.float32(1.0) / np.float32(b * m), np.float32)
    w1, w2 = oracle.chamfer_backward(x, y, gd1, gd2, i1, i2)
    for got, ref in ((g1.numpy(), w1), (g2.numpy(), w2)):
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    assert isinstance(pkg.chamfer_loss_step_host(hx, hy), float)


def test_host_buffer_loss_pipeline(pkg, oracle, cuda):
    """ChamferLossPipeline (psd_chamfer_loss_step_host_ex, sync=0): double-buffered steps return each step's own loss."""
    b, n, m = 2, 512, 768
    steps = []
    for seed in range(5):
        x, y = make_clouds("uniform", b, n, m, seed=100 + seed)
        d1, d2, _, _ = oracle.chamfer_forward(x, y, nthreads=4)
        steps.append((torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory(),
                      d1.astype(np.float64).mean() + d2.astype(np.float64).mean()))
    pipe = pkg.ChamferLossPipeline(cuda)
    got = []
    for s, (hx, hy, _) in enumerate(steps):
        pipe.submit(hx, hy)
        if s > 0:
            got.append(pipe.result())
    got.append(pipe.result())
    for g, (_, _, want) in zip(got, steps):
        assert abs(g - want) <= 1e-5 * abs(want)


@pytest.mark.parametrize("depth", [2, 3, 4])
def test_host_buffer_loss_pipeline_graph_replay(pkg, oracle, cuda, depth):
    """The same pinned staging buffers submitted again and again are replayed from a cached CUDA graph
    (psd_host_step_graphs): the loss must follow the buffers' CURRENT contents and agree with the plain path."""
    from importlib import import_module
    lib = import_module(pkg.__name__ + "._lib").lib
    b, n, m = 2, 640, 512
    bufs = [(torch.empty(b, n, 3).pin_memory(), torch.empty(b, m, 3).pin_memory()) for _ in range(depth)]
    data = [make_clouds("uniform", b, n, m, seed=300 + s) for s in range(4 * depth)]
    want = []
    for x, y in data:
        d1, d2, _, _ = oracle.chamfer_forward(x, y, nthreads=4)
        want.append(d1.astype(np.float64).mean() + d2.astype(np.float64).mean())
    results = {}
    for graphs in (1, 0):
        old = lib.psd_host_step_graphs(graphs)
        try:
            pipe = pkg.ChamferLossPipeline(cuda, depth=depth)
            got = []
            for s, (x, y) in enumerate(data):
                if s >= depth:
                    got.append(pipe.result())   # the step that used this buffer pair has finished: safe to overwrite
                hx, hy = bufs[s % depth]
                hx.copy_(torch.from_numpy(x)); hy.copy_(torch.from_numpy(y))
                pipe.submit(hx, hy)
            while pipe.pending:
                got.append(pipe.result())
        finally:
            lib.psd_host_step_graphs(old if old in (0, 1) else 1)
        results[graphs] = got
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert abs(g - w) <= 1e-5 * abs(w)
    for g0, g1 in zip(results[0], results[1]):   # the per-cloud sums are float atomics: equal up to summation order
        assert abs(g0 - g1) <= 1e-6 * abs(g0)


@pytest.mark.parametrize("ctas", [1, 37, 74, 0])
def test_tensor_core_kernel_with_capped_grid(pkg, oracle, cuda, ctas):
    """psd_chamfer_tc_ctas: a launch limited to a part of the SMs (for callers that keep several launches in flight) gives the
    same bits; two capped launches on two streams at once as well."""
    lib = pkg._lib.lib
    x, y = make_clouds("uniform", 24, 1024, 1100, seed=77)
    want = oracle.chamfer_forward(x, y, nthreads=8)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    oldv, oldc = lib.psd_chamfer_nn_variant(3), lib.psd_chamfer_tc_ctas(ctas)
    try:
        outs = []
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                outs.append(pkg.chamfer_3DDist()(tx, ty))
        torch.cuda.synchronize()
    finally:
        lib.psd_chamfer_nn_variant(oldv); lib.psd_chamfer_tc_ctas(oldc)
    for out in outs:
        assert_bit_equal([t.cpu().numpy() for t in out], want, f"ctas={ctas}")


def test_backward_raw_accumulates_into_given_buffers(pkg, oracle, cuda):
    """chamfer_3D.backward adds onto the caller's buffers (the reference relies on caller-zeroed grads)."""
    x, y = make_clouds("uniform", 2, 256, 300, seed=21)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    d1, d2, i1, i2 = pkg.chamfer_3DDist()(tx, ty)
    g1 = torch.ones_like(d1)
    g2 = torch.ones_like(d2)
    a1 = torch.full_like(tx, 0.5)
    a2 = torch.full_like(ty, -0.25)
    assert pkg.chamfer_3D.backward(tx, ty, a1, a2, g1, g2, i1, i2) == 1
    w1, w2 = oracle.chamfer_backward(x, y, g1.cpu().numpy(), g2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy())
    assert np.allclose(a1.cpu().numpy(), w1 + 0.5, rtol=1e-5, atol=1e-6)
    assert np.allclose(a2.cpu().numpy(), w2 - 0.25, rtol=1e-5, atol=1e-6)


def test_fused_fscore_and_sums(pkg, oracle, cuda):
    x, y = make_clouds("clustered", 4, 1500, 1100, seed=31)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    out = pkg.chamfer_fscore_fused(tx, ty, threshold=1e-4)
    want = oracle.chamfer_forward(x, y, nthreads=8)
    assert_bit_equal([out[k].cpu().numpy() for k in ("dist1", "dist2", "idx1", "idx2")], want, "fused")
    c1, c2 = oracle.fscore_counts(want[0], want[1], 1e-4)
    assert (out["counts"].cpu().numpy() == np.stack([c1, c2], 1)).all()
    sums = out["sums"].cpu().numpy()
    assert np.allclose(sums[:, 0], want[0].astype(np.float64).sum(1), rtol=1e-5)
    assert np.allclose(sums[:, 1], want[1].astype(np.float64).sum(1), rtol=1e-5)
    f, p1, p2 = pkg.fscore(out["dist1"], out["dist2"], 1e-4)
    assert torch.equal(f, out["fscore"]) and torch.equal(p1, out["precision_1"]) and torch.equal(p2, out["precision_2"])


def test_soa_layout_matches_aos(pkg, oracle, cuda):
    """[B,3,N] input (the generator's native layout, train.py:163) gives the same bits without a transpose copy."""
    x, y = make_clouds("uniform", 3, 777, 1029, seed=41)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    out = pkg.chamfer_fscore_fused(tx.transpose(1, 2).contiguous(), ty.transpose(1, 2).contiguous(), layout=3)
    want = oracle.chamfer_forward(x, y, nthreads=8)
    assert_bit_equal([out[k].cpu().numpy() for k in ("dist1", "dist2", "idx1", "idx2")], want, "soa")


def test_query_slices_cover_whole(pkg, oracle, cuda):
    """Query sharding: two launches over disjoint query slices reproduce the full result."""
    x, y = make_clouds("uniform", 2, 1000, 1300, seed=51)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    d1 = torch.zeros(2, 1000, device=cuda); d2 = torch.zeros(2, 1300, device=cuda)
    i1 = torch.zeros(2, 1000, device=cuda, dtype=torch.int32); i2 = torch.zeros(2, 1300, device=cuda, dtype=torch.int32)
    # slices are expressed per direction through q_begin/q_count on the larger cloud; use halves of max(n, m)
    for qb, qc in ((0, 650), (650, 650)):
        assert pkg.chamfer_3D.forward_ex(tx, ty, d1, d2, i1, i2, q_begin=qb, q_count=qc) == 1
    want = oracle.chamfer_forward(x, y, nthreads=8)
    assert_bit_equal([t.cpu().numpy() for t in (d1, d2, i1, i2)], want, "slices")


def test_host_buffer_entry_point(pkg, oracle, cuda):
    import ctypes
    x, y = make_clouds("uniform", 2, 512, 640, seed=61)
    d1 = np.zeros((2, 512), np.float32); d2 = np.zeros((2, 640), np.float32)
    i1 = np.zeros((2, 512), np.int32); i2 = np.zeros((2, 640), np.int32)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    rc = pkg._lib.lib.psd_chamfer_forward_host(vp(x), vp(y), 2, 512, 640, vp(d1), vp(d2), vp(i1), vp(i2), None)
    assert rc == 1, pkg._lib.last_error()
    assert_bit_equal((d1, d2, i1, i2), oracle.chamfer_forward(x, y), "host entry")


def test_errors_are_loud(pkg, cuda):
    x = torch.rand(2, 64, 3)
    with pytest.raises(RuntimeError):
        pkg.chamfer_3DDist()(x, x)  # CPU tensors: no fallback
    xd = torch.rand(2, 64, 3, device=cuda, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        pkg.chamfer_3DDist()(xd, xd)


def test_query_sharding_with_cuda_ops_emulated_ranks(pkg, oracle, cuda):
    """The CUDA local operator of sharding.py, two 'ranks' run one after the other in this process (no
    process group: the all_reduce is emulated by summing the disjoint slices)."""
    from importlib import import_module
    import psd_b200
    sh = import_module(psd_b200.PKG_NAME + ".sharding")
    x, y = make_clouds("uniform", 2, 1500, 900, seed=71)
    tx, ty = torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)
    outs = [sh.chamfer_query_sharded(tx, ty, r, 2, threshold=1e-4, assemble=False) for r in range(2)]
    want = oracle.chamfer_forward(x, y, nthreads=8)
    got = [sum(o[k] for o in outs).cpu().numpy() for k in ("dist1", "dist2", "idx1", "idx2")]
    assert_bit_equal(got, want, "query-sharded")
    c1, c2 = oracle.fscore_counts(want[0], want[1], 1e-4)
    assert (sum(o["counts"] for o in outs).cpu().numpy() == np.stack([c1, c2], 1)).all()
    g1 = torch.rand(2, 1500, device=cuda); g2 = torch.rand(2, 900, device=cuda)
    i1 = torch.from_numpy(want[2]).to(cuda); i2 = torch.from_numpy(want[3]).to(cuda)
    parts = [sh.chamfer_backward_query_sharded(tx, ty, g1, g2, i1, i2, r, 2) for r in range(2)]
    w1, w2 = oracle.chamfer_backward(x, y, g1.cpu().numpy(), g2.cpu().numpy(), want[2], want[3])
    assert np.allclose((parts[0][0] + parts[1][0]).cpu().numpy(), w1, rtol=1e-5, atol=1e-6)
    assert np.allclose((parts[0][1] + parts[1][1]).cpu().numpy(), w2, rtol=1e-5, atol=1e-6)


def test_large_cloud_properties(pkg, oracle, cuda):
    """BASELINE.json configs[4] scale (N=M=131072 is too slow for the CPU oracle): size-independent properties --
    dist equals the exact distance along idx, no sampled target is closer, F-score counts match a recount,
    and a 16k-query slice is bit-exact against the oracle."""
    b, n = 1, 131072
    g = torch.Generator().manual_seed(5)
    x = torch.rand(b, n, 3, generator=g)
    y = torch.rand(b, n, 3, generator=g)
    tx, ty = x.to(cuda), y.to(cuda)
    out = pkg.chamfer_fscore_fused(tx, ty, threshold=1e-4)
    d1, i1 = out["dist1"], out["idx1"].long()
    sel = torch.gather(ty, 1, i1.unsqueeze(-1).expand(-1, -1, 3))
    diff = sel - tx
    dd = torch.addcmul(torch.addcmul(diff[..., 1] * diff[..., 1], diff[..., 0], diff[..., 0]), diff[..., 2], diff[..., 2])
    assert torch.allclose(dd, d1, rtol=1e-6, atol=0)
    probe = ty[:, torch.randint(0, n, (64,), generator=g)]
    pd = ((tx[:, :, None, :] - probe[:, None, :, :]) ** 2).sum(-1).min(2)[0]
    assert bool((d1 <= pd * (1 + 1e-5)).all())
    assert int(out["counts"][0, 0]) == int((d1 < 1e-4).sum()) and int(out["counts"][0, 1]) == int((out["dist2"] < 1e-4).sum())
    # bit-exact slice against the oracle: first 2048 queries of direction 1 against all targets
    wd, _, wi, _ = oracle.chamfer_forward(x[:, :2048].numpy(), y.numpy(), nthreads=16)
    assert np.array_equal(d1[:, :2048].cpu().numpy(), wd) and np.array_equal(out["idx1"][:, :2048].cpu().numpy(), wi)


def test_stream_order_with_programmatic_dependent_launch(pkg, oracle, cuda):
    """The tensor-core forward kernel is launched as a programmatic dependent of whatever precedes it on the stream (PSD_PDL): its
    set-up may overlap the predecessor's drain, its first global access must not.  (1) inputs PRODUCED by a kernel that was
    launched right in front of it, no synchronisation in between; (2) back-to-back forward / backward / forward launches on
    different inputs that REUSE the same output and gradient buffers through the raw entry points.  Everything bit-equal to the
    oracle, every repetition."""
    b, n, m = 24, 1024, 1024                      # >= 2 units per SM: the automatic choice takes the tensor-core kernel as well
    sets = [make_clouds(kind, b, n, m, seed=70 + i) for i, kind in enumerate(("uniform", "clustered", "uniform"))]
    wants = [oracle.chamfer_forward(x, y, nthreads=8) for x, y in sets]
    bases = [(torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda)) for x, y in sets]
    torch.cuda.synchronize()
    # (1) producer kernels directly in front of the launch
    for rep in range(4):
        for (bx, by), want in zip(bases, wants):
            tx = bx * 2.0 - bx          # elementwise kernels on the same stream: tx == bx bit for bit, written just now
            ty = by * 2.0 - by
            got = pkg.chamfer_3DDist()(tx, ty)
            assert_bit_equal([g.cpu().numpy() for g in got], want, f"producer in front, repetition {rep}")
    # (2) raw entry points, shared output buffers, forward -> backward -> forward without synchronisation
    d1 = torch.empty(b, n, device=cuda); d2 = torch.empty(b, m, device=cuda)
    i1 = torch.empty(b, n, device=cuda, dtype=torch.int32); i2 = torch.empty(b, m, device=cuda, dtype=torch.int32)
    g1 = torch.zeros(b, n, 3, device=cuda); g2 = torch.zeros(b, m, 3, device=cuda)
    gd1 = torch.ones(b, n, device=cuda); gd2 = torch.ones(b, m, device=cuda)
    snaps = []
    for rep in range(3):
        for k, (bx, by) in enumerate(bases):
            assert pkg.chamfer_3D.forward(bx, by, d1, d2, i1, i2) == 1
            g1.zero_(); g2.zero_()
            assert pkg.chamfer_3D.backward(bx, by, g1, g2, gd1, gd2, i1, i2) == 1
            snaps.append((k, d1.clone(), d2.clone(), i1.clone(), i2.clone(), g1.clone(), g2.clone()))
    torch.cuda.synchronize()
    ref_grads = {}
    for k, sd1, sd2, si1, si2, sg1, sg2 in snaps:
        assert_bit_equal([sd1.cpu().numpy(), sd2.cpu().numpy(), si1.cpu().numpy(), si2.cpu().numpy()], wants[k], f"back-to-back set {k}")
        if k not in ref_grads:
            x, y = sets[k]
            ref_grads[k] = oracle.chamfer_backward(x, y, np.ones((b, n), np.float32), np.ones((b, m), np.float32), wants[k][2], wants[k][3])
        for got, want in zip((sg1, sg2), ref_grads[k]):
            got = got.cpu().numpy()
            assert np.allclose(got, want, rtol=1e-5, atol=1e-6), f"gradient of set {k}: max abs diff {np.abs(got - want).max()}"
