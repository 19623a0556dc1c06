"""Quick device-side timing of the hot path next to the reference's CUDA extensions (oracle/_ref), CUDA events."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import psd_b200

pkg = psd_b200.load()
dev = torch.device("cuda:0")


def timeit(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def load_ref():
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_chamfer_3D"))
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_emd"))
        import ref_chamfer_3D, ref_emd
        return ref_chamfer_3D, ref_emd
    except Exception as e:  # noqa
        print("reference extensions unavailable:", e)
        return None, None


def main():
    B, N = 32, 2048
    if len(sys.argv) > 2 and sys.argv[1].isdigit():
        B, N = int(sys.argv[1]), int(sys.argv[2])
    torch.manual_seed(0)
    x = torch.rand(B, N, 3).to(dev); y = torch.rand(B, N, 3).to(dev)
    d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, N, device=dev)
    i1 = torch.empty(B, N, device=dev, dtype=torch.int32); i2 = torch.empty(B, N, device=dev, dtype=torch.int32)
    g1 = torch.rand(B, N, device=dev); g2 = torch.rand(B, N, device=dev)
    gx = torch.zeros(2 * B * N * 3, device=dev)
    gx1 = gx[: B * N * 3].view(B, N, 3); gx2 = gx[B * N * 3:].view(B, N, 3)
    pairs = 2.0 * B * N * N
    fwd = lambda: pkg.chamfer_3D.forward(x, y, d1, d2, i1, i2)
    def bwd():
        gx.zero_()
        pkg.chamfer_3D.backward(x, y, gx1, gx2, g1, g2, i1, i2)
    def both():
        fwd(); bwd()
    peak = np.zeros(1, np.float32)
    import ctypes
    tf = ctypes.c_float(0)
    pkg._lib.lib.psd_fp32_fma_peak(ctypes.c_float(1.0), ctypes.byref(tf), None)
    print(f"measured FP32 FMA peak: {tf.value:.2f} TFLOP/s")
    def many(fn, k=20):
        def f():
            for _ in range(k):
                fn()
        return f
    for name, fn, k in (("chamfer fwd x20", many(fwd), 20), ("chamfer bwd x20", many(lambda: pkg.chamfer_3D.backward(x, y, gx1, gx2, g1, g2, i1, i2)), 20)):
        best, med = timeit(fn, reps=10, warm=2)
        print(f"{name:24s} per launch: best {best / k:8.2f} us  median {med / k:8.2f} us   {8 * pairs / (med / k * 1e-6) / 1e12:.2f} TFLOP/s-equivalent")
    for name, fn in (("chamfer fwd", fwd), ("chamfer bwd(+zero)", bwd), ("chamfer fwd+bwd", both)):
        best, med = timeit(fn)
        print(f"{name:24s} best {best:8.1f} us  median {med:8.1f} us   {pairs / (med * 1e-6) / 1e12:.3f} Tpairs/s  "
              f"{8 * pairs / (med * 1e-6) / 1e12:.2f} TFLOP/s (8 flop/pair)")
    # CUDA graph replay of fwd+bwd
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        both(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            both()
    best, med = timeit(lambda: g.replay())
    print(f"{'fwd+bwd (graph replay)':24s} best {best:8.1f} us  median {med:8.1f} us   {8 * pairs / (med * 1e-6) / 1e12:.2f} TFLOP/s")
    fb = np.zeros(2, np.int64)
    pkg._lib.lib.psd_chamfer_stats(fb.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 0)
    print("fallback queries total:", fb[1])

    if "--no-emd" in sys.argv:
        return
    # EMD
    n = N
    dist = torch.empty(B, n, device=dev); ass = torch.empty(B, n, device=dev, dtype=torch.int32)
    emdf = lambda: pkg.emd.forward_fresh(x, y, dist, ass, 0.005, 50)
    best, med = timeit(emdf, reps=10, warm=2)
    print(f"{'emd fwd (eps .005, 50)':24s} best {best:8.1f} us  median {med:8.1f} us   {B / (med * 1e-6):.0f} clouds/s")
    for cs in (1, 2, 4, 8):
        inv = torch.full((B, n), -1, device=dev, dtype=torch.int32); price = torch.zeros(B, n, device=dev)
        def f():
            ass.fill_(-1); inv.fill_(-1); price.zero_()
            pkg._lib.lib.psd_emd_forward_cluster(pkg._lib.ptr(x), pkg._lib.ptr(y), B, n, pkg._lib.ptr(dist), pkg._lib.ptr(ass),
                                                 pkg._lib.ptr(price), pkg._lib.ptr(inv), ctypes.c_float(0.005), 50, cs, None)
        best, med = timeit(f, reps=5, warm=1)
        print(f"  emd cluster={cs}: median {med:8.1f} us")

    rc, re = load_ref()
    if rc is not None:
        rd1 = torch.zeros(B, N, device=dev); rd2 = torch.zeros(B, N, device=dev)
        ri1 = torch.zeros(B, N, device=dev, dtype=torch.int32); ri2 = torch.zeros(B, N, device=dev, dtype=torch.int32)
        rfwd = lambda: rc.forward(x, y, rd1, rd2, ri1, ri2)
        def rbwd():
            gx.zero_(); rc.backward(x, y, gx1, gx2, g1, g2, ri1, ri2)
        def rboth():
            rfwd(); rbwd()
        for name, fn in (("REF chamfer fwd", rfwd), ("REF chamfer bwd(+zero)", rbwd), ("REF chamfer fwd+bwd", rboth)):
            best, med = timeit(fn)
            print(f"{name:24s} best {best:8.1f} us  median {med:8.1f} us   {8 * pairs / (med * 1e-6) / 1e12:.2f} TFLOP/s")
        print("ref vs ours idx equal:", bool((ri1 == i1).all() and (ri2 == i2).all()), " dist equal:", bool(torch.equal(rd1, d1) and torch.equal(rd2, d2)))
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        def remd():
            rdist = z(B, n); rass = z(B, n, dt=torch.int32) - 1; rinv = z(B, n, dt=torch.int32) - 1; rprice = z(B, n)
            rbid = z(B, n, dt=torch.int32); rbi = z(B, n); rmi = z(B, n); ru = z(B * n, dt=torch.int32); rmx = z(B * n, dt=torch.int32)
            c1 = z(512, dt=torch.int32); c2 = z(512, dt=torch.int32); c3 = z(512, dt=torch.int32)
            re.forward(x, y, rdist, rass, rprice, rinv, rbid, rbi, rmi, ru, c1, c2, c3, rmx, 0.005, 50)
            return rdist, rass
        best, med = timeit(remd, reps=5, warm=1)
        print(f"{'REF emd fwd':24s} best {best:8.1f} us  median {med:8.1f} us   {B / (med * 1e-6):.0f} clouds/s")
        rdist, rass = remd(); emdf(); torch.cuda.synchronize()
        print("ref vs ours emd assignment equal:", bool((rass == ass).all()), " mismatching clouds:", int(((rass != ass).any(1)).sum()),
              " loss rel diff:", float((rdist.sqrt().mean() - dist.sqrt().mean()).abs() / rdist.sqrt().mean()))


if __name__ == "__main__":
    main()
