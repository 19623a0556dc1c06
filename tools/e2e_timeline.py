"""GPU timeline (CUPTI via torch.profiler) of a few pipelined host training steps at config 2."""
import importlib, json, os, sys, time
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
lib = importlib.import_module("3d-pointcloudreconstruction_b200._lib").lib
B, N, M = 32, 2048, 2048
torch.manual_seed(0)
hxy = [torch.rand(B * (N + M), 3).pin_memory() for _ in range(4)]
hv = [(h[: B * N].view(B, N, 3), h[B * N:].view(B, M, 3)) for h in hxy]
dev = torch.device("cuda", 0)
lib.psd_host_step_graphs(int(os.environ.get("GRAPHS", "0")))
pipe = pkg.ChamferLossPipeline(dev, depth=int(os.environ.get('DEPTH', '4')))
def run(k):
    for s in range(k):
        if len(pipe.pending) == pipe.depth: pipe.result()
        pipe.submit(*hv[s % 4])
    while pipe.pending: pipe.result()
run(50)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(16)
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "e2e_trace.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset", "cuda_runtime")]
gpu = sorted((e for e in ev if e["cat"] != "cuda_runtime"), key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
for e in gpu:
    print(f"{e['ts'] - t0:9.1f} +{e['dur']:7.1f}  stream {e['args'].get('stream')}  {e['cat']:10s} {e['name'][:60]}")
rt = sorted((e for e in ev if e["cat"] == "cuda_runtime"), key=lambda e: e["ts"])
print("---- runtime")
for e in rt[:60]:
    print(f"{e['ts'] - t0:9.1f} +{e['dur']:7.1f}  {e['name'][:50]}")
