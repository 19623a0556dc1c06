"""A/B of the chamfer backward forms at config 2 (B=32, N=M=2048), graph replay, CUDA events, L2-resident and HBM-cold:
  accumulate : torch zero fill + psd_chamfer_backward (all terms atomic; the reference's contract)
  fused zero : psd_chamfer_forward_zero (the forward launch zero-fills the gradients) + psd_chamfer_backward
  overwrite  : psd_chamfer_backward_ex(overwrite=1): store launch + scatter launch
and the forward + backward step in the serial form and with 8 chains x 37-CTA forward launches in flight."""
import ctypes, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import psd_b200

pkg = psd_b200.load(); lib = pkg._lib.lib
dev = torch.device("cuda:0")
B, N = 32, 2048
pool = 64
g = torch.Generator().manual_seed(0)
xs = torch.rand(pool, B, N, 3, generator=g).to(dev); ys = torch.rand(pool, B, N, 3, generator=g).to(dev)
d1 = torch.empty(pool, B, N, device=dev); d2 = torch.empty(pool, B, N, device=dev)
i1 = torch.empty(pool, B, N, device=dev, dtype=torch.int32); i2 = torch.empty(pool, B, N, device=dev, dtype=torch.int32)
gd1 = torch.rand(pool, B, N, generator=g).to(dev); gd2 = torch.rand(pool, B, N, generator=g).to(dev)
gb = torch.empty(pool, 6 * B * N, device=dev)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
cur = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(64 * 1024 * 1024, device=dev)

def fwd(p): assert pkg.chamfer_3D.forward(xs[p], ys[p], d1[p], d2[p], i1[p], i2[p]) == 1
def bwd_acc(p):
    gb[p].zero_()
    assert lib.psd_chamfer_backward(vp(xs[p]), vp(ys[p]), vp(gb[p][:3*B*N]), vp(gb[p][3*B*N:]), vp(gd1[p]), vp(gd2[p]), vp(i1[p]), vp(i2[p]), B, N, N, cur()) == 1
def fwd_zero(p):
    assert lib.psd_chamfer_forward_zero(vp(xs[p]), vp(ys[p]), B, N, N, 0, vp(d1[p]), vp(d2[p]), vp(i1[p]), vp(i2[p]), None, 0.0, None, vp(gb[p]), gb[p].numel(), cur()) == 1
def bwd_acc_nozero(p):
    assert lib.psd_chamfer_backward(vp(xs[p]), vp(ys[p]), vp(gb[p][:3*B*N]), vp(gb[p][3*B*N:]), vp(gd1[p]), vp(gd2[p]), vp(i1[p]), vp(i2[p]), B, N, N, cur()) == 1
def bwd_ow(p):
    assert lib.psd_chamfer_backward_ex(vp(xs[p]), vp(ys[p]), vp(gb[p][:3*B*N]), vp(gb[p][3*B*N:]), vp(gd1[p]), vp(gd2[p]), vp(i1[p]), vp(i2[p]), B, N, N, 0, 1, cur()) == 1

stream = torch.cuda.Stream()
def timed(body, reps=48, chains=1, cold=True):
    with torch.cuda.stream(stream):
        for p in range(4): body(p)
        stream.synchronize()
        sides = [torch.cuda.Stream() for _ in range(chains - 1)]
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            for sd in sides: sd.wait_stream(stream)
            for s in range(reps):
                c = s % chains
                if c:
                    with torch.cuda.stream(sides[c - 1]): body((5 + s) % pool)
                else: body((5 + s) % pool)
            for sd in sides: stream.wait_stream(sd)
        gr.replay(); stream.synchronize()
        ts = []
        for _ in range(7):
            if cold: flush.fill_(1.0)
            stream.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); gr.replay(); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return statistics.median(ts)

for p in range(pool): fwd(p)
torch.cuda.synchronize()
print(f"forward alone (serial)            : {timed(fwd):7.2f} us")
print(f"forward + fused zero fill         : {timed(fwd_zero):7.2f} us")
for name, fn in (("accumulate (zero fill + atomics)", bwd_acc), ("accumulate, pre-zeroed buffers", bwd_acc_nozero), ("overwrite (two launches)", bwd_ow)):
    print(f"backward {name:32s}: {timed(fn):7.2f} us cold, {timed(fn, cold=False):7.2f} us warm")
for name, f, bk in (("memset + accumulate", fwd, bwd_acc), ("fused zero + accumulate", fwd_zero, bwd_acc_nozero), ("overwrite", fwd, bwd_ow)):
    step = lambda p, f=f, bk=bk: (f(p), bk(p))
    lib.psd_chamfer_tc_ctas(0)
    ser = timed(step)
    lib.psd_chamfer_tc_ctas(37)
    pip = timed(step, reps=48, chains=8)
    lib.psd_chamfer_tc_ctas(0)
    print(f"step fwd+bwd, {name:24s}: serial {ser:7.2f} us, 8 chains x 37 CTAs {pip:7.2f} us")
