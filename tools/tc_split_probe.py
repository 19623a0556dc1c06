"""Two chamfer forward launches in flight on two streams, each limited to a part of the SMs (psd_chamfer_tc_ctas), against one
launch at a time on all SMs: throughput at config 2."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
lib = pkg._lib.lib
dev = torch.device("cuda", 0)
B, N = 32, 2048
g = torch.Generator().manual_seed(0)
xs = [torch.rand(B, N, 3, generator=g).to(dev) for _ in range(8)]; ys = [torch.rand(B, N, 3, generator=g).to(dev) for _ in range(8)]
outs = [(torch.empty(B, N, device=dev), torch.empty(B, N, device=dev), torch.empty(B, N, device=dev, dtype=torch.int32), torch.empty(B, N, device=dev, dtype=torch.int32)) for _ in range(8)]
streams = [torch.cuda.Stream() for _ in range(8)]
def run(nstreams, reps=200):
    for r in range(reps):
        s = r % nstreams
        with torch.cuda.stream(streams[s]):
            pkg.chamfer_3D.forward(xs[s], ys[s], *outs[s])
def timed(nstreams, reps=200):
    run(nstreams, 20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(nstreams, reps)
    for s in streams[:nstreams]: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
ref = [o.clone() for o in outs[0]]
for ctas, ns in ((0, 1), (0, 2), (74, 2), (74, 1), (49, 3), (37, 4), (29, 5), (24, 6), (18, 8), (37, 8), (0, 4)):
    lib.psd_chamfer_tc_ctas(ctas)
    us = timed(ns)
    same = all(torch.equal(a, b) for a, b in zip(outs[0], ref)) if ctas else True
    if ctas == 0 and ns == 1: ref = [o.clone() for o in outs[0]]
    print(f"max_ctas={ctas:3d} streams={ns}: {us:6.2f} us per forward   identical={same}")
lib.psd_chamfer_tc_ctas(0)
