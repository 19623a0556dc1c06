"""Timings of the ops around the hot path: farthest point sampling, cont_proj splat, proj min-dist; CUDA events,
oracle port on the host for scale."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
from oracle import oracle
dev = torch.device("cuda", 0)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
rng = np.random.default_rng(0)
for b, n, npoint in ((32, 1024, 128), (32, 1024, 256), (148, 2048, 256), (32, 16384, 1024)):
    x = rng.random((b, n, 3), dtype=np.float32); t = torch.from_numpy(x).to(dev)
    us = timed(lambda: pkg.utils.farthest_point_sample(t, npoint))
    t0 = time.perf_counter(); oracle.farthest_point_sample(x[:2], npoint); cpu = (time.perf_counter() - t0) / 2 * b
    print(f"fps   B={b} N={n} npoint={npoint}: {us:9.1f} us ({us / npoint:.2f} us per round); numpy port {cpu * 1e3:.0f} ms per batch")
for b, n, h, w in ((32, 1024, 64, 64), (128, 1024, 64, 64), (32, 1024, 128, 128)):
    p = (rng.random((b, n, 3), dtype=np.float32) * 2 - 1); t = torch.from_numpy(p).to(dev)
    us = timed(lambda: pkg.projection.cont_proj(t, h, w, dev, 0.5))
    t0 = time.perf_counter(); oracle.cont_proj(p[:2], h, w, 0.5); cpu = (time.perf_counter() - t0) / 2 * b
    print(f"splat B={b} N={n} {h}x{w}: {us:9.1f} us; numpy port {cpu * 1e3:.0f} ms per batch")
    img = pkg.projection.cont_proj(t, h, w, dev, 0.5).clamp(0, 1)
    dm = torch.from_numpy(oracle.grid_dist(h, w).astype(np.float32) + 1)
    for mode in ("as_written", "intended"):
        us = timed(lambda: pkg.proj_loss.min_dist_terms(img, img, dm, mode=mode), reps=5)
        print(f"proj  B={b} {h}x{w} {mode}: {us:9.1f} us (incl. table H2D)")
