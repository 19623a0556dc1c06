#!/usr/bin/env python
"""bench.py -- headline benchmark of the point-set-distance hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    chamfer3D forward + backward, B=32 clouds per GPU, N=M=2048 points, fp32, U[0,1)^3 synthetic clouds.
    One "step" = one forward (dist1, dist2, idx1, idx2) + zeroing of the gradients + one backward
    (grad_xyz1, grad_xyz2) over one batch.  metric = directed point pairs per second = 2*B*N*M / t(step),
    whole job (all ranks).  Weak scaling: every rank owns its own batch of B clouds, no data-path collective.

`value`      : batches resident in HBM, K steps replayed as one CUDA graph, CUDA events on the launch stream.
               Every step uses a different batch from a pool whose footprint exceeds the 126 MB L2
               ("inputs larger than L2"), so inputs are read from HBM.
`e2e`        : the same metric through the public API (chamfer_3DDist()(xyz1, xyz2) + .backward()), with the
               step's inputs copied from pinned host memory and the loss read back on the host every step.
`roofline`   : dominant kernel chamfer_nn_tc_kernel, algorithmic 8 flop per directed pair,
               duration = CUDA events around back-to-back launches, peak = FP32 FMA rate measured live by an
               FFMA-only kernel (MEASURED_PEAKS.json has no FP32 entry; nominal 74.45 TFLOP/s also given).
               The kernel evaluates the pairs as a split-fp16 K=16 GEMM on the tensor cores (tcgen05) and is bound
               by the min-reduction on the ALU pipe and the TMEM hand-shake, so `roofline.tensor` adds the issued
               tensor flops against MEASURED_PEAKS.json's bf16 figure.
`cpu_baseline`: the oracle's C restatement of the same step on the host cores (bounded sample).
--impl reference: the reference's own CPU implementation of the path (its pure-torch chamfer,
               loss/loss_.py:66-91, restated in oracle/oracle.py) on the host cores, bounded sample per step.
Extra keys: `emd` (BASELINE configs[2], clouds/s), `reference_cuda` (the reference's CUDA extensions built
unmodified into oracle/_ref, timed on the same GPU: the "same box" bar of the north star).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, N, M = 32, 2048, 2048
EMD_EPS, EMD_ITERS = 0.005, 50
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45
TRAFFIC_BYTES = 1.6e6   # dram bytes per chamfer forward launch, from the committed ncu capture (profiles/)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip emd / reference_cuda / cpu_baseline extras")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def device_index_for_nvml(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """CPU arm: the reference's pure-torch chamfer (fp64 expansion) forward + autograd backward."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    xa = torch.rand(4, N, 3)
    ya = torch.rand(4, M, 3)
    steps, warm = max(1, args.steps), max(0, args.warmup)

    def make_step(bs):
        x, y = xa[:bs], ya[:bs]

        def step():
            a = x.clone().requires_grad_(True)
            b = y.clone().requires_grad_(True)
            P = O.torch_batched_pairwise_dist(a, b)
            loss = torch.min(P, 2)[0].float().mean() + torch.min(P, 1)[0].float().mean()
            loss.backward()
            return float(loss)
        return step

    # bounded sample: 4, 2 or 1 of the 32 clouds per step, sized so that the K + W steps asked for end within ~2.5 minutes
    bs = 4
    step = make_step(bs)
    step()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    while bs > 1 and t1 * (steps + warm) > 150.0:
        bs //= 2
        t1 *= 0.5
    step = make_step(bs)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    pairs = 2.0 * bs * N * M
    v = pairs / dt
    sample = f"{bs} of {B} clouds per step (N=M={N}), {steps} steps, torch fp64 xx+yy-2*bmm + autograd backward"
    emit({
        "impl": "reference", "metric": "chamfer_fwd_bwd_point_pairs_per_s", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"chamfer3D fwd+bwd B={B} N=M={N} fp32 (BASELINE configs[1])", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# --------------------------------------------------------------------------------------------------
def event_time_ms(torch, fn, stream=None):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else a library prints there (NCCL's version banner, for
    one) is sent to stderr instead.  The JSON line goes to the saved descriptor through emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import psd_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    pkg = psd_b200.load()
    L = pkg._lib
    K, W = args.steps, max(args.warmup, 3)

    # ---- pool of batches larger than L2: every step reads a batch that is not cache resident
    per_batch = 4 * (3 * B * N + 3 * B * M) + 8 * (B * N + B * M) + 4 * (3 * B * N + 3 * B * M) + 4 * (B * N + B * M)
    pool = max(2, int(2.2 * 126e6 / per_batch) + 1)
    g = torch.Generator().manual_seed(1234 + rank)
    xs = torch.rand(pool, B, N, 3, generator=g).to(dev)
    ys = torch.rand(pool, B, M, 3, generator=g).to(dev)
    d1 = torch.empty(pool, B, N, device=dev); d2 = torch.empty(pool, B, M, device=dev)
    i1 = torch.empty(pool, B, N, device=dev, dtype=torch.int32); i2 = torch.empty(pool, B, M, device=dev, dtype=torch.int32)
    gd1 = torch.rand(pool, B, N, generator=g).to(dev); gd2 = torch.rand(pool, B, M, generator=g).to(dev)
    gbuf = torch.empty(pool, 3 * B * (N + M), device=dev)

    def step(s):
        p = s % pool
        assert pkg.chamfer_3D.forward(xs[p], ys[p], d1[p], d2[p], i1[p], i2[p]) == 1, L.last_error()
        gbuf[p].zero_()
        g1 = gbuf[p][: 3 * B * N].view(B, N, 3)
        g2 = gbuf[p][3 * B * N:].view(B, M, 3)
        assert pkg.chamfer_3D.backward(xs[p], ys[p], g1, g2, gd1[p], gd2[p], i1[p], i2[p]) == 1, L.last_error()

    CHAINS = int(os.environ.get("PSD_BENCH_CHAINS", "8"))
    TC_CTAS = int(os.environ.get("PSD_BENCH_TC_CTAS", "0")) or max(1, torch.cuda.get_device_properties(dev).multi_processor_count // 4)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for s in range(W):  # untimed warm-up (also loads the module before graph capture)
            step(s)
        stream.synchronize()
        # one launch at a time on all SMs: the serial picture of a step (reported as config.serial_ms_per_step)
        L.lib.psd_chamfer_tc_ctas(0)
        gser = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gser, stream=stream):
            for s in range(50):
                step(W + s)
        gser.replay(); stream.synchronize()
        serial_ms = min(event_time_ms(torch, gser.replay, stream) for _ in range(3)) / 50
        # Steps are independent batches, so CHAINS (8) of them are kept in flight (as in the pipelined host loop): step s is
        # captured on chain s % CHAINS, and every launch of the tensor-core NN kernel is limited to a quarter of the SMs
        # (psd_chamfer_tc_ctas(37)).  A CTA then owns four times as many units, so its serial prologue and tail amortise, and
        # with twice as many launches in flight as SM quarters a finished CTA's SM is taken over at once by a waiting launch
        # (tools/tc_split_probe.py: 37.0 us per forward with one launch at a time, 26.8 us with eight 37-CTA launches in
        # flight; sweep of chains x cap in profiles/r1_chamfer_nn_tc_summary.md).  Every step still runs its full forward +
        # zero + backward.
        L.lib.psd_chamfer_tc_ctas(TC_CTAS)
        sides = [torch.cuda.Stream(device=dev) for _ in range(CHAINS - 1)]
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for sd in sides:
                sd.wait_stream(stream)
            for s in range(K):
                c = s % CHAINS
                if c:
                    with torch.cuda.stream(sides[c - 1]):
                        step(W + s)
                else:
                    step(W + s)
            for sd in sides:
                stream.wait_stream(sd)
        graph.replay()  # one untimed replay
        stream.synchronize()

    sampler = ClockSampler(device_index_for_nvml(local_rank))
    sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ms = event_time_ms(torch, graph.replay, stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    # keep the sampler running over a few more replays so that it sees the load (the timed region is short)
    with torch.cuda.stream(stream):
        for _ in range(3):
            graph.replay()
        stream.synchronize()
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    pairs_step = 2.0 * B * N * M
    value = world * pairs_step * K / (ms_total * 1e-3)

    # ---- end to end through the public API with host buffers
    # End to end through the C ABI's host-buffer entry point of the path's caller, Loss.get_chamfer_loss + backward
    # (loss/loss.py:30-37): psd_chamfer_loss_step_host copies the step's clouds from pinned host memory, runs forward,
    # fused mean loss and backward, and returns the loss on the host (gradients stay on the device for the caller's
    # own backward).  Both clouds of a step live in one pinned buffer [B*(N+M), 3] -> one H2D copy per step.
    NBUF = DEPTH = 8   # host staging buffers = steps in flight in the pipelined e2e loop
    hxy = [torch.rand(B * (N + M), 3, generator=g).pin_memory() for _ in range(NBUF)]
    hviews = [(h[: B * N].view(B, N, 3), h[B * N:].view(B, M, 3)) for h in hxy]

    def e2e_step(s):
        a, b_ = hviews[s % NBUF]
        return pkg.chamfer_loss_step_host(a, b_)   # H2D + fwd + loss + bwd + D2H(loss) + sync

    pipe = pkg.ChamferLossPipeline(dev, depth=DEPTH)

    def e2e_run_pipelined(nsteps):
        """The same steps pipelined (psd_chamfer_loss_step_host_ex, sync=0, DEPTH steps in flight): the H2D copy of a step
        and the host's latency between submits overlap the kernels of the steps before it; every step's loss is still
        read on the host."""
        acc = 0.0
        for s in range(nsteps):
            a, b_ = hviews[s % NBUF]
            if len(pipe.pending) == pipe.depth:
                acc += pipe.result()
            pipe.submit(a, b_)
        while pipe.pending:
            acc += pipe.result()
        return acc

    def e2e_step_torch(s):   # the same step through the torch-facing module API (reported as e2e.torch_api)
        xy = hxy[s % NBUF].to(dev, non_blocking=True)
        a = xy[: B * N].view(B, N, 3).requires_grad_(True)
        b_ = xy[B * N:].view(B, M, 3).requires_grad_(True)
        loss = loss_mod.get_chamfer_loss(a, b_)
        loss.backward()
        return loss.item()

    L.lib.psd_chamfer_tc_ctas(0)   # one step at a time in the torch-API and blocking legs: all SMs per launch
    loss_mod = pkg.Loss()
    for s in range(W):
        e2e_step_torch(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(50):
        e2e_step_torch(s)
    torch.cuda.synchronize()
    e2e_torch_value = world * 2.0 * B * N * M * 50 / (time.perf_counter() - t0)

    Ke = min(K, 400)
    for s in range(W):
        e2e_step(s)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(Ke):
        e2e_step(s)
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3
    L.lib.psd_chamfer_tc_ctas(TC_CTAS)   # DEPTH steps in flight: the launches share the SMs (captured into the step graphs)
    e2e_run_pipelined(W + 2 * DEPTH)   # every (buffer, slot) combination seen twice: its CUDA graph is cached
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_run_pipelined(Ke)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    te = torch.tensor([e2e_ms, e2e_sync_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * pairs_step * Ke / (float(te[0].item()) * 1e-3)
    e2e_sync_value = world * pairs_step * Ke / (float(te[1].item()) * 1e-3)

    # ---- EMD (BASELINE configs[2]) on every GPU of the job: each rank its own B clouds (weak scaling), max over ranks
    emd_all = None
    if not args.no_extras:
        ex_, ey_ = xs[0], ys[0]
        ed_ = torch.empty(B, N, device=dev); ea_ = torch.empty(B, N, device=dev, dtype=torch.int32)
        for _ in range(2):
            pkg.emd.forward_fresh(ex_, ey_, ed_, ea_, EMD_EPS, EMD_ITERS)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        reps_e = 10
        ems = event_time_ms(torch, lambda: [pkg.emd.forward_fresh(ex_, ey_, ed_, ea_, EMD_EPS, EMD_ITERS) for _ in range(reps_e)]) / reps_e
        tm = torch.tensor([ems], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        emd_all = {"clouds_per_s": world * B / (float(tm.item()) * 1e-3), "ms": float(tm.item()), "n_gpus": world,
                   "config": f"B={B} per GPU, n={N}, eps={EMD_EPS}, iters={EMD_ITERS}; {reps_e} back-to-back launches, max over ranks"}

    out = {
        "metric": "chamfer_fwd_bwd_point_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"chamfer3D fwd+bwd B={B} per GPU, N=M={N}, fp32, bit-exact idx (BASELINE configs[1])",
                   "cache": f"inputs larger than L2: {pool} batches x {per_batch / 1e6:.1f} MB rotate, one per step",
                   "timing": f"K steps captured in one CUDA graph as {CHAINS} independent chains (step s on chain s % {CHAINS}), every tensor-core NN launch limited to {TC_CTAS} CTAs so that the launches in flight share the SMs; CUDA events on the launch stream, max over ranks",
                   "serial_ms_per_step": serial_ms, "serial_note": "the same step with one launch at a time on all SMs",
                   "parallelism": f"batch-sharded x{world}, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 4 * 3 * B * (N + M), "d2h_bytes_per_step": 4,
                "steps": Ke, "pipeline_depth": DEPTH, "tc_ctas_per_launch": TC_CTAS,
                "api": "psd_chamfer_loss_step_host_ex (C ABI, pinned host buffers), pipelined: per step H2D + chamfer fwd + mean loss [loss/loss.py:36] + bwd + D2H loss; up to 8 steps in flight on 8 streams/workspaces, each replayed from a cached CUDA graph, so the H2D copies (1.57 MB = 32 us at ~49 GB/s PCIe, the bound) and the host latency overlap the kernels; every step's loss is read on the host",
                "synchronous": {"value": e2e_sync_value, "unit": "pairs/s", "api": "psd_chamfer_loss_step_host: the same step, one blocking call per step (no overlap)"},
                "torch_api": {"value": e2e_torch_value, "unit": "pairs/s", "api": "Loss().get_chamfer_loss(pred, gt); loss.backward(); loss.item() with a pinned-host H2D copy per step"}},
        "gpu_launches": 2 * K,   # value leg: chamfer_nn_tc_kernel + chamfer_grad_kernel per step (e2e adds chamfer_mean_loss_kernel)
        "clocks": sampler.result(),
    }
    if emd_all is not None:
        out["emd_all_gpus"] = emd_all

    L.lib.psd_chamfer_tc_ctas(0)
    if rank == 0:
        # ---- roofline of the dominant kernel: chamfer_nn_kernel alone (one launch at a time, all SMs), back-to-back launches, CUDA events
        reps = 50
        with torch.cuda.stream(stream):
            for s in range(5):
                pkg.chamfer_3D.forward(xs[s % pool], ys[s % pool], d1[s % pool], d2[s % pool], i1[s % pool], i2[s % pool])
            stream.synchronize()
            g2_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2_, stream=stream):
                for s in range(reps):
                    p = (7 + s) % pool
                    pkg.chamfer_3D.forward(xs[p], ys[p], d1[p], d2[p], i1[p], i2[p])
            g2_.replay(); stream.synchronize()
            fwd_ms = min(event_time_ms(torch, g2_.replay, stream) for _ in range(3)) / reps
            g3_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g3_, stream=stream):
                for s in range(reps):
                    p = (7 + s) % pool
                    pkg.chamfer_3D.backward(xs[p], ys[p], gbuf[p][: 3 * B * N].view(B, N, 3), gbuf[p][3 * B * N:].view(B, M, 3),
                                            gd1[p], gd2[p], i1[p], i2[p])
            g3_.replay(); stream.synchronize()
            bwd_ms = min(event_time_ms(torch, g3_.replay, stream) for _ in range(3)) / reps
        tf = ctypes.c_float(0)
        L.lib.psd_fp32_fma_peak(ctypes.c_float(1.0), ctypes.byref(tf), None)
        torch.cuda.synchronize()
        achieved = 8.0 * pairs_step / (fwd_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        issued_tensor = 2.0 * 16 * pairs_step / (fwd_ms * 1e-3) / 1e12   # K = 16 MACs per pair on the tensor pipe
        tensor_peak = float(peaks.get("bf16_tflops", 1590.0))
        out["roofline"] = {
            "bound": "fp32_fma", "kernel": "chamfer_nn_tc_kernel", "achieved": achieved, "peak": float(tf.value),
            "unit": "TFLOP/s", "frac": achieved / float(tf.value), "peak_source": "measured live: FFMA-only kernel on all SMs (psd_fp32_fma_peak)",
            "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS,
            "algorithmic": "8 flop per directed pair x 2*B*N*M pairs per launch", "us_per_launch": fwd_ms * 1e3,
            "traffic": TRAFFIC_BYTES, "traffic_note": "dram__bytes_read+write per launch from profiles/ (ncu --set full); inputs 1.57 MB",
            "note": "pair work runs as a split-fp16 GEMM on tcgen05 tensor cores; 8 flop/pair is the reference formulation's count, "
                    "so the fraction is 'FP32-FMA-equivalent' and may exceed what an FFMA kernel can reach (the FFMA kernel of this "
                    "library: 0.53)",
            "tensor": {"issued_tflops": issued_tensor, "peak": tensor_peak, "frac": issued_tensor / tensor_peak,
                       "what": "32 fp16 flop per pair (K=16) issued with tcgen05.mma kind::f16; peak = MEASURED_PEAKS.json bf16_tflops"},
        }
        bwd_bytes = 4.0 * (3 * B * (N + M)) + 8.0 * B * (N + M) + 4.0 * (3 * B * (N + M))
        out["roofline_bwd"] = {
            "bound": "hbm", "kernel": "chamfer_grad_kernel", "achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "frac": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / hbm_peak, "us_per_launch": bwd_ms * 1e3,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s", "traffic": None,
            "note": "4.2 MB per launch: latency/atomic bound, not bandwidth bound",
        }
        fbv = np.zeros(2, np.int64)
        L.lib.psd_chamfer_stats(fbv.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 0)
        out["config"]["exact_fallback_queries_total"] = int(fbv[1])

        if not args.no_extras:
            # ---- EMD (BASELINE configs[2]) on this GPU
            ex, ey = xs[0], ys[0]
            edist = torch.empty(B, N, device=dev); eass = torch.empty(B, N, device=dev, dtype=torch.int32)
            for _ in range(2):
                pkg.emd.forward_fresh(ex, ey, edist, eass, EMD_EPS, EMD_ITERS)
            torch.cuda.synchronize()
            emd_ms = min(event_time_ms(torch, lambda: pkg.emd.forward_fresh(ex, ey, edist, eass, EMD_EPS, EMD_ITERS)) for _ in range(5))
            out["emd"] = {"clouds_per_s": B / (emd_ms * 1e-3), "ms": emd_ms, "config": f"B={B} n={N} eps={EMD_EPS} iters={EMD_ITERS}",
                          "launches": 1, "scope": "rank 0's GPU; emd_all_gpus is the whole job"}
            # the training setting of the same op (loss/loss.py:18-28: eps=0.05, iters=3000, generator output n=1024)
            tx_, ty_ = ex[:, :1024].contiguous(), ey[:, :1024].contiguous()
            tdist = torch.empty(B, 1024, device=dev); tass = torch.empty(B, 1024, device=dev, dtype=torch.int32)
            pkg.emd.forward_fresh(tx_, ty_, tdist, tass, 0.05, 3000)
            torch.cuda.synchronize()
            emd_train_ms = min(event_time_ms(torch, lambda: pkg.emd.forward_fresh(tx_, ty_, tdist, tass, 0.05, 3000)) for _ in range(3))
            out["emd_train"] = {"clouds_per_s": B / (emd_train_ms * 1e-3), "ms": emd_train_ms,
                                "config": f"B={B} n=1024 eps=0.05 iters=3000 (Loss.get_emd_loss)", "launches": 1}
            # ---- the reference's CUDA extensions on the same GPU (oracle/_ref, built unmodified)
            try:
                sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_chamfer_3D"))
                sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_emd"))
                import ref_chamfer_3D
                import ref_emd
                rd1 = torch.zeros(B, N, device=dev); rd2 = torch.zeros(B, M, device=dev)
                ri1 = torch.zeros(B, N, device=dev, dtype=torch.int32); ri2 = torch.zeros(B, M, device=dev, dtype=torch.int32)
                rg = torch.zeros(3 * B * (N + M), device=dev)

                def ref_step(p):
                    ref_chamfer_3D.forward(xs[p], ys[p], rd1, rd2, ri1, ri2)
                    rg.zero_()
                    ref_chamfer_3D.backward(xs[p], ys[p], rg[: 3 * B * N].view(B, N, 3), rg[3 * B * N:].view(B, M, 3), gd1[p], gd2[p], ri1, ri2)
                for s in range(3):
                    ref_step(s)
                torch.cuda.synchronize()
                rms = event_time_ms(torch, lambda: [ref_step((11 + s) % pool) for s in range(20)]) / 20
                z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)

                def ref_emd_step():
                    a = [z(B, N), z(B, N, dt=torch.int32) - 1, z(B, N), z(B, N, dt=torch.int32) - 1, z(B, N, dt=torch.int32), z(B, N), z(B, N),
                         z(B * N, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(B * N, dt=torch.int32)]
                    ref_emd.forward(ex, ey, *a, EMD_EPS, EMD_ITERS)
                ref_emd_step(); torch.cuda.synchronize()
                rems = min(event_time_ms(torch, ref_emd_step) for _ in range(3))
                def ref_emd_train_step():
                    n1 = 1024
                    a = [z(B, n1), z(B, n1, dt=torch.int32) - 1, z(B, n1), z(B, n1, dt=torch.int32) - 1, z(B, n1, dt=torch.int32), z(B, n1), z(B, n1),
                         z(B * n1, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(B * n1, dt=torch.int32)]
                    ref_emd.forward(tx_, ty_, *a, 0.05, 3000)
                ref_emd_train_step(); torch.cuda.synchronize()
                retms = event_time_ms(torch, ref_emd_train_step)
                out["reference_cuda"] = {"chamfer_fwd_bwd_pairs_per_s": pairs_step / (rms * 1e-3), "chamfer_ms_per_step": rms,
                                         "emd_clouds_per_s": B / (rems * 1e-3), "emd_ms": rems,
                                         "emd_train_clouds_per_s": B / (retms * 1e-3), "emd_train_ms": retms,
                                         "what": "reference chamfer3D/emd extensions compiled unmodified for sm_100a (oracle/_ref), same GPU, default stream, CUDA events"}
            except Exception as e:  # noqa: BLE001
                out["reference_cuda"] = {"unavailable": repr(e)[:200]}
            # ---- CPU baseline: the oracle's C restatement, bounded sample, all host cores
            try:
                from oracle import oracle as O
                cores = os.cpu_count() or 1
                bs = 8
                cx = xs[0][:bs].cpu().numpy(); cy = ys[0][:bs].cpu().numpy()
                cg1 = gd1[0][:bs].cpu().numpy(); cg2 = gd2[0][:bs].cpu().numpy()
                O.chamfer_forward(cx[:1], cy[:1], nthreads=cores)
                t0 = time.perf_counter()
                reps_c = 3
                for _ in range(reps_c):
                    c = O.chamfer_forward(cx, cy, nthreads=cores)
                    O.chamfer_backward(cx, cy, cg1, cg2, c[2], c[3])
                dtc = (time.perf_counter() - t0) / reps_c
                out["cpu_baseline"] = {"value": 2.0 * bs * N * M / dtc, "unit": "pairs/s", "cores": cores, "kind": "port",
                                       "sample": f"{bs} of {B} clouds (N=M={N}) fwd+bwd x{reps_c}, oracle/psd_oracle.c with OpenMP over clouds"}
            except Exception as e:  # noqa: BLE001
                out["cpu_baseline"] = {"unavailable": repr(e)[:200]}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
