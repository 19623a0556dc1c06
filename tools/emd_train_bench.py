"""EMD at the training setting (loss/loss.py:18-28: eps=0.05, iters=3000, generator output n=1024, B=32): this library's
single persistent launch against the reference extension's 7 launches per iteration, same GPU, same inputs."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
dev = torch.device("cuda", 0)
def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_emd"))
try:
    import ref_emd
except Exception as e:
    ref_emd = None; print("reference extension unavailable:", e)
for B, N, eps, iters in ((32, 1024, 0.05, 3000), (32, 2048, 0.05, 3000), (32, 2048, 0.005, 50)):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, N, 3, generator=g).to(dev); y = torch.rand(B, N, 3, generator=g).to(dev)
    d = torch.empty(B, N, device=dev); a = torch.empty(B, N, device=dev, dtype=torch.int32)
    ms = ev(lambda: pkg.emd.forward_fresh(x, y, d, a, eps, iters), 5)
    line = f"B={B} n={N} eps={eps} iters={iters}: ours {ms:8.3f} ms ({B / ms * 1e3:8.0f} clouds/s)"
    if ref_emd is not None:
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        out = {}
        def ref():
            st = [z(B, N), z(B, N, dt=torch.int32) - 1, z(B, N), z(B, N, dt=torch.int32) - 1, z(B, N, dt=torch.int32), z(B, N), z(B, N),
                  z(B * N, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(B * N, dt=torch.int32)]
            ref_emd.forward(x, y, *st, eps, iters); out["d"], out["a"] = st[0], st[1]
        rms = ev(ref, 2)
        same = bool((out["a"] == a).all()); nd = int((out["a"] != a).any(1).sum())
        line += f"   reference {rms:9.3f} ms ({B / rms * 1e3:7.0f} clouds/s)  speed-up {rms / ms:6.1f}x  assignment equal: {same} ({nd} clouds differ)"
    print(line)
