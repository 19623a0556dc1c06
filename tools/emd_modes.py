"""A/B of the auction kernel's switches (object grid, solo mode) at BASELINE config 3 and at the training setting."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
lib = pkg._lib.lib
dev = torch.device("cuda", 0)
def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for B, N, eps, iters in ((32, 2048, 0.005, 50), (32, 1024, 0.05, 3000), (64, 2048, 0.005, 50), (8, 2048, 0.005, 50), (32, 4096, 0.005, 50)):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, N, 3, generator=g).to(dev); y = torch.rand(B, N, 3, generator=g).to(dev)
    d = torch.empty(B, N, device=dev); a = torch.empty(B, N, device=dev, dtype=torch.int32)
    out = []
    for grid in (0, 1):
        for solo in (0, 1):
            lib.psd_emd_grid_mode(grid); lib.psd_emd_solo_mode(solo)
            out.append(f"grid={grid} solo={solo}: {ev(lambda: pkg.emd.forward_fresh(x, y, d, a, eps, iters)):7.3f} ms")
    lib.psd_emd_grid_mode(1); lib.psd_emd_solo_mode(1)
    print(f"B={B} n={N} eps={eps} iters={iters}:  " + "   ".join(out))
