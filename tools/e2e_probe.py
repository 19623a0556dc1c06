"""Host-side cost of the end-to-end training step (psd_chamfer_loss_step_host_ex) at config 2, with and without the cached
CUDA graph: CPU time of submit(), steady-state time per pipelined step."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3d-pointcloudreconstruction_b200")
lib = importlib.import_module("3d-pointcloudreconstruction_b200._lib").lib
B, N, M = 32, 2048, 2048
torch.manual_seed(0)
hxy = [torch.rand(B * (N + M), 3).pin_memory() for _ in range(4)]
hv = [(h[: B * N].view(B, N, 3), h[B * N:].view(B, M, 3)) for h in hxy]
dev = torch.device("cuda", 0)
for graphs, depth in ((1, 2), (1, 3), (0, 4), (1, 4), (1, 5), (1, 6), (1, 8), (0, 8)):
    lib.psd_host_step_graphs(graphs)
    pipe = pkg.ChamferLossPipeline(dev, depth=depth)
    def run(k):
        acc = 0.0; tsub = 0.0
        for s in range(k):
            a, b_ = hv[s % 4]
            if len(pipe.pending) == pipe.depth: acc += pipe.result()
            t0 = time.perf_counter(); pipe.submit(a, b_); tsub += time.perf_counter() - t0
        while pipe.pending: acc += pipe.result()
        return tsub
    run(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tsub = run(400); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"graphs={graphs} depth={depth}: {dt / 400 * 1e6:7.1f} us/step  submit() {tsub / 400 * 1e6:6.1f} us  -> {2.0 * B * N * M * 400 / dt:.3e} pairs/s")
