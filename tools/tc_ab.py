"""A/B of builds of libpsd_b200.so (same ABI, different kernel organisation): forward of config 2 (B=32, N=M=2048), serial form and 8
chains x 37 CTAs, graph replay + CUDA events, and a checksum of dist/idx so that every build can be seen to give the same bits.

    python tools/tc_ab.py                          # the in-tree build and every exp/libpsd_*.so
    PSD_B200_LIB=... python tools/tc_ab.py --one   # (internal) time the library the environment names
"""
import glob, hashlib, os, statistics, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--one" not in sys.argv:
    libs = [None] + sorted(glob.glob(os.path.join(ROOT, "exp", "libpsd_*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["PSD_B200_LIB"] = lib
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
        print(f"{os.path.basename(lib) if lib else 'in-tree build':24s} {out.stdout.strip() or out.stderr[-400:]}")
    sys.exit(0)

import torch
import psd_b200
pkg = psd_b200.load(); lib = pkg._lib.lib
dev = torch.device("cuda:0")
B, N, pool = 32, 2048, 64
g = torch.Generator().manual_seed(0)
xs = torch.rand(pool, B, N, 3, generator=g).to(dev); ys = torch.rand(pool, B, N, 3, generator=g).to(dev)
d1 = torch.empty(pool, B, N, device=dev); d2 = torch.empty(pool, B, N, device=dev)
i1 = torch.empty(pool, B, N, device=dev, dtype=torch.int32); i2 = torch.empty(pool, B, N, device=dev, dtype=torch.int32)
flush = torch.empty(64 * 1024 * 1024, device=dev)
def fwd(p): assert pkg.chamfer_3D.forward(xs[p], ys[p], d1[p], d2[p], i1[p], i2[p]) == 1
stream = torch.cuda.Stream()
def timed(reps=48, chains=1):
    with torch.cuda.stream(stream):
        for p in range(4): fwd(p)
        stream.synchronize()
        sides = [torch.cuda.Stream() for _ in range(chains - 1)]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            for sd in sides: sd.wait_stream(stream)
            for s in range(reps):
                c = s % chains
                if c:
                    with torch.cuda.stream(sides[c - 1]): fwd((5 + s) % pool)
                else: fwd((5 + s) % pool)
            for sd in sides: stream.wait_stream(sd)
        gr.replay(); stream.synchronize()
        ts = []
        for _ in range(9):
            flush.fill_(1.0); stream.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); gr.replay(); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return statistics.median(ts)
lib.psd_chamfer_tc_ctas(0); ser = timed()
lib.psd_chamfer_tc_ctas(37); pip = timed(chains=8)
lib.psd_chamfer_tc_ctas(0)
for p in range(pool): fwd(p)
torch.cuda.synchronize()
h = hashlib.sha1()
for t in (d1, d2, i1, i2): h.update(t.cpu().numpy().tobytes())
import ctypes, numpy as np
fb = np.zeros(2, np.int64); lib.psd_chamfer_stats(fb.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 0)
print(f"serial {ser:6.2f} us   8x37 {pip:6.2f} us   sha1 {h.hexdigest()[:12]}   exact-scan fallbacks {int(fb[1])}")
