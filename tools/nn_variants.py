"""Time the two chamfer NN forward kernels (psd_chamfer_nn_variant 1 = shared-block, 2 = grouped) against each
other on a few shapes and check that their outputs are bit-identical.  CUDA-graph replay of `reps` launches over
a rotating pool of batches, CUDA events on the launch stream.

    python tools/nn_variants.py [B N M]...      default: 32x2048x2048, 32x1024x1024, 8x8192x8192, 2x131072x131072
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import psd_b200

pkg = psd_b200.load()
L = pkg._lib.lib
dev = torch.device("cuda:0")
VARIANTS = [int(v) for v in os.environ.get("NN_VARIANTS", "1,3").split(",")]


def time_variant(variant, xs, ys, outs, reps):
    L.psd_chamfer_nn_variant(variant)
    pool = len(xs)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for s in range(2):
            assert pkg.chamfer_3D.forward(xs[s % pool], ys[s % pool], *outs[s % pool]) == 1
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for s in range(reps):
                pkg.chamfer_3D.forward(xs[s % pool], ys[s % pool], *outs[s % pool])
        g.replay()
        stream.synchronize()
        best = 1e30
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); g.replay(); e1.record(stream); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
    L.psd_chamfer_nn_variant(0)
    return best * 1e3  # us


def main():
    shapes = [(32, 2048, 2048), (32, 1024, 1024), (64, 2048, 2048), (8, 8192, 8192), (2, 131072, 131072)]
    args = [int(a) for a in sys.argv[1:]]
    if args:
        shapes = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)]
    for (b, n, m) in shapes:
        pairs = 2.0 * b * n * m
        big = pairs > 1e10
        pool = 2 if big else 8
        reps = 2 if big else 20
        g = torch.Generator().manual_seed(7)
        xs = [torch.rand(b, n, 3, generator=g).to(dev) for _ in range(pool)]
        ys = [torch.rand(b, m, 3, generator=g).to(dev) for _ in range(pool)]
        res = {}
        for v in VARIANTS:
            outs = [(torch.empty(b, n, device=dev), torch.empty(b, m, device=dev),
                     torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, m, device=dev, dtype=torch.int32))
                    for _ in range(pool)]
            us = time_variant(v, xs, ys, outs, reps)
            torch.cuda.synchronize()
            res[v] = (us, outs)
        same = all(torch.equal(a, b_) for oa, ob in zip(res[VARIANTS[0]][1], res[VARIANTS[-1]][1]) for a, b_ in zip(oa, ob))
        line = f"B={b} N={n} M={m}: "
        for v in VARIANTS:
            name = {1: "shared-block", 2: "grouped", 3: "tensor-core"}[v]
            us = res[v][0]
            line += f"{name} {us:9.1f} us ({8 * pairs / (us * 1e-6) / 1e12:5.1f} TFLOP/s-alg)  "
        print(line + f"identical={same}", flush=True)


if __name__ == "__main__":
    main()
