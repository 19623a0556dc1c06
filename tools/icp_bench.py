"""Batched ICP (psd_icp_batch) at the eval setting (testnet.py:62-64: per sample icp(points, fake, tolerance=1e-10,
max_iterations=1024), N=1024) against the numpy/brute-force oracle port on the host."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
from oracle import oracle
def rot(axis, th):
    axis = np.asarray(axis, dtype=np.float64); axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
for batch, n in ((32, 1024), (148, 1024), (32, 2048), (32, 4096)):
    rng = np.random.default_rng(0)
    B = rng.random((batch, n, 3)).astype(np.float32); A = np.empty_like(B)
    for s in range(batch):
        A[s] = ((B[s].astype(np.float64) - 0.5) @ rot(rng.standard_normal(3), 0.2).T + 0.5 + 0.02 + 0.02 * rng.standard_normal((n, 3))).astype(np.float32)[rng.permutation(n)]
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for _ in range(2): T, d, it = pkg.icp.icp_batch(a, b, max_iterations=1024, tolerance=1e-10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): T, d, it = pkg.icp.icp_batch(a, b, max_iterations=1024, tolerance=1e-10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    its = it.cpu().numpy() + 1
    t0 = time.perf_counter(); wT, wd, wi = oracle.icp(A[0], B[0], max_iterations=1024, tolerance=1e-10); cpu = time.perf_counter() - t0
    print(f"batch={batch} n={n}: {ms:8.3f} ms per batch, iterations/sample mean {its.mean():.1f} max {its.max()}; "
          f"{ms * 1e3 / its.max():.1f} us per iteration of the slowest sample; oracle port 1 sample {cpu * 1e3:.0f} ms "
          f"(T diff {np.abs(T[0].cpu().numpy() - wT).max():.1e})")
