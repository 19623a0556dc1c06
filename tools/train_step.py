#!/usr/bin/env python
"""Synthetic 3D-FENet train step (BASELINE.json configs[3]): the hot path's share of a real optimizer step.

The reference's training loop (train.py:148-177) is re-created as a benchmark harness, not as a component: a random-init
generator of the reference's architecture (models/repvgg_edge_nose_NEW_cmlp.py:253-336: RepVGG-A2 encoder, Laplacian edge branch,
three-scale MLP / Conv1d point decoder emitting [B,3,1024]) in plain torch, synthetic images, Adam as train.py:115, and the
loss schedule of epochs 1-30, 100*CD + 100*EMD (train.py:162-165), plus the projection terms of finetune.py:154-165
(silhouettes of the detached clouds, BCE + min-distance terms, logged only -- they carry no gradient in the reference either).
The edge branch's Linear is sized 3*(H/4)*(W/4) so that 224x224 images work (the reference hard-wires 128x128).
Data-parallel over the job's ranks with DistributedDataParallel (NCCL gradient all-reduce, ~176 M fp32 parameters).

    python tools/train_step.py [--ops ours|reference] [--steps K] [--batch B] [--image 224]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py ...

`--ops reference` swaps ONLY the two native modules: the reference's unmodified wrappers (oracle/_ref/py) over its own CUDA
extensions compiled for sm_100a (oracle/_ref) -- everything else in the step is identical.  Per-phase device times come from
CUDA events; the exposed gradient all-reduce is the difference between a normal step and a `no_sync()` step."""
import argparse
import importlib
import importlib.util
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ---------------------------------------------------------------------------------------------- load generator (plain torch)
class _RepBlock(nn.Module):
    """Training-form RepVGG block: 3x3 conv+BN, 1x1 conv+BN and (same shape only) a BN identity branch, summed, ReLU."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.k3 = nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout))
        self.k1 = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, 0, bias=False), nn.BatchNorm2d(cout))
        self.skip = nn.BatchNorm2d(cin) if (cin == cout and stride == 1) else None

    def forward(self, x):
        y = self.k3(x) + self.k1(x)
        if self.skip is not None:
            y = y + self.skip(x)
        return F.relu(y)


def _stage(cin, cout, depth):
    return nn.Sequential(*[_RepBlock(cin if i == 0 else cout, cout, 2 if i == 0 else 1) for i in range(depth)])


class Generator(nn.Module):
    """RepVGG-A2 widths (64, 96, 192, 384, 1408), depths (1, 2, 4, 14, 1); 1000-d image code + 1000-d edge code -> three clouds."""

    def __init__(self, image=224):
        super().__init__()
        widths, depths = (64, 96, 192, 384, 1408), (1, 2, 4, 14, 1)
        chans = (3,) + widths
        self.encoder = nn.Sequential(*[_stage(chans[i], chans[i + 1], depths[i]) for i in range(5)])
        self.code = nn.Linear(widths[-1], 1000)
        lap = torch.full((3, 3, 3, 3), -1.0 / 3.0)
        lap[:, :, 1, 1] = 8.0 / 3.0
        self.register_buffer("laplace", lap)
        self.edge = nn.Sequential(nn.Conv2d(3, 16, 3, 2, 1, bias=False), nn.BatchNorm2d(16), nn.ReLU(inplace=True),
                                  nn.Conv2d(16, 3, 3, 2, 1, bias=False), nn.BatchNorm2d(3), nn.ReLU(inplace=True))
        self.edge_code = nn.Linear(3 * (image // 4) * (image // 4), 1000)
        self.trunk = nn.ModuleList([nn.Linear(2000, 1024), nn.Linear(1024, 512), nn.Linear(512, 256)])
        self.coarse = nn.Linear(256, 128 * 3)
        self.mid = nn.Linear(512, 128 * 128)
        self.mid_conv = nn.Conv1d(128, 6, 1)
        self.fine = nn.Linear(1024, 256 * 512)
        self.fine_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.ReLU(), nn.Conv1d(512, 256, 1), nn.ReLU(), nn.Conv1d(256, 12, 1))

    def forward(self, img):
        b = img.shape[0]
        e = self.edge_code(self.edge(F.conv2d(img, self.laplace, padding=1)).flatten(1))
        z = self.code(self.encoder(img).mean((2, 3)))
        f1 = F.relu(self.trunk[0](torch.cat([z, e], 1)))
        f2 = F.relu(self.trunk[1](f1))
        f3 = F.relu(self.trunk[2](f2))
        p1 = self.coarse(f3).view(b, 128, 1, 3)
        p2 = p1 + self.mid_conv(F.relu(self.mid(f2)).view(b, 128, 128)).transpose(1, 2).reshape(b, 128, 2, 3)
        p2 = p2.reshape(b, 256, 1, 3)
        p3 = p2 + self.fine_conv(F.relu(self.fine(f1)).view(b, 512, 256)).transpose(1, 2).reshape(b, 256, 4, 3)
        # the reference returns [B,3,N] tensors (its callers transpose them back, train.py:163)
        return (p1.reshape(b, 128, 3).transpose(1, 2).contiguous(), p2.reshape(b, 256, 3).transpose(1, 2).contiguous(),
                p3.reshape(b, 1024, 3).transpose(1, 2).contiguous())


# ---------------------------------------------------------------------------------------------- the two op sets
def load_ops(kind):
    """Returns (chamfer_loss(pred_view, gt), emd_loss(pred_view, gt), label)."""
    import psd_b200
    pkg = psd_b200.load()
    if kind == "ours":
        loss = pkg.Loss()
        return loss.get_chamfer_loss, loss.get_emd_loss, pkg
    # the reference's wrappers over the reference's extensions (module names aliased to what the wrappers import)
    ref = os.path.join(ROOT, "oracle", "_ref")
    for d in ("ref_chamfer_3D", "ref_emd"):
        sys.path.insert(0, os.path.join(ref, d))
    import ref_chamfer_3D
    import ref_emd
    sys.modules["chamfer_3D"] = ref_chamfer_3D
    sys.modules["emd"] = ref_emd
    importlib.find_loader = lambda name: (sys.modules.get(name) or importlib.util.find_spec(name))   # py3.12 shim
    for name in ("dist_chamfer_3D", "emd_module"):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(ref, "py", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules["ref_" + name] = mod
    cham = sys.modules["ref_dist_chamfer_3D"].chamfer_3DDist()
    emdm = sys.modules["ref_emd_module"].emdModule()

    def chamfer_loss(pred, gt):            # loss/loss.py:30-37
        d1, d2, _, _ = cham(pred, gt)
        return torch.mean(d1) + torch.mean(d2)

    def emd_loss(pred, gt):                # loss/loss.py:18-28
        d, _ = emdm(pred, gt, eps=0.05, iters=3000)
        return torch.sqrt(d).mean(1).mean()
    return chamfer_loss, emd_loss, pkg


class Timer:
    def __init__(self):
        self.ev = {}

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.setdefault(name, []).append(e)


def run(ops="ours", steps=10, warmup=3, batch=32, image=224, proj=True, quiet=False):
    """One process of the job (rank from the environment).  Returns the rank-0 report dict (None elsewhere)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_pg = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
        own_pg = True
    torch.manual_seed(1234)                                     # identical initial weights on every rank
    chamfer_loss, emd_loss, pkg = load_ops(ops)
    gen = Generator(image).to(dev)
    nparam = sum(p.numel() for p in gen.parameters())
    model = gen
    if world > 1:
        model = nn.parallel.DistributedDataParallel(gen, device_ids=[local], gradient_as_bucket_view=True)
    opt = torch.optim.Adam(gen.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4)   # train.py:115
    g = torch.Generator().manual_seed(77 + rank)
    images = torch.randn(batch, 3, image, image, generator=g).pin_memory()
    points = torch.rand(batch, 1024, 3, generator=g).pin_memory()
    # the [64,64,64,64] pixel-distance matrix of finetune.py:154 lives on the device here (a 67 MB host tensor that is cloned
    # and incremented every step would stall the host, not measure the path)
    dist_mat = torch.from_numpy(pkg.proj_loss.grid_dist(64, 64)).float().to(dev) if proj else None

    phases = ("h2d", "model_fwd", "chamfer_fwd", "emd_fwd", "proj", "backward", "optimizer")

    def step(tm, sync_grads=True):
        tm.mark("start")
        img = images.to(dev, non_blocking=True)
        pts = points.to(dev, non_blocking=True)
        tm.mark("h2d")
        _, _, fake = model(img)                                  # [B,3,1024]
        tm.mark("model_fwd")
        pred = fake.transpose(2, 1)                              # train.py:163: a VIEW of the generator's output
        cd = chamfer_loss(pred, pts)
        tm.mark("chamfer_fwd")
        em = emd_loss(pred, pts)
        tm.mark("emd_fwd")
        logs = None
        if proj:                                                 # finetune.py:154-165: detached clouds, logged terms only
            with torch.no_grad():
                pp = pkg.projection.cont_proj(pred.detach().contiguous() * 2 - 1, 64, 64, dev, 0.5).clamp(0, 1)
                pg = pkg.projection.cont_proj(pts * 2 - 1, 64, 64, dev, 0.5).clamp(0, 1)
                dm = dist_mat.clone()
                bce, fwd_d, bwd_d = pkg.proj_loss.get_loss_proj(pp, pg, dev, "bce_prob", 1.0, True, dm)
                logs = (bce, fwd_d.mean(), bwd_d.mean())
        tm.mark("proj")
        total = 100.0 * cd + 100.0 * em                          # train.py:162-165 (epochs 1-30)
        opt.zero_grad(set_to_none=True)
        if world > 1 and not sync_grads:
            with model.no_sync():
                total.backward()
        else:
            total.backward()
        tm.mark("backward")
        opt.step()
        tm.mark("optimizer")
        return total, logs

    first_loss = None
    for i in range(warmup):
        l0, _ = step(Timer())
        if i == 0:
            first_loss = float(l0.detach())      # same weights and data for every op set: a parity check of the two op sets
    torch.cuda.synchronize()

    def timed(nsteps, sync_grads):
        tms = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = None
        for _ in range(nsteps):
            tm = Timer()
            last, _ = step(tm, sync_grads)
            tms.append(tm)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / nsteps
        if world > 1:
            dist.barrier()
        per = {}
        for ph_prev, ph in zip(("start",) + phases[:-1], phases):
            per[ph] = sum(t.ev[ph_prev][0].elapsed_time(t.ev[ph][0]) for t in tms) / nsteps
        return wall * 1e3, per, float(last.detach())

    wall_ms, per, loss_v = timed(steps, True)
    wall_nosync_ms = None
    if world > 1:
        wall_nosync_ms, _, _ = timed(max(3, steps // 2), False)
    stats = torch.tensor([wall_ms, wall_nosync_ms or 0.0] + [per[p] for p in phases], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    stats = stats.tolist()
    report = None
    if rank == 0:
        wall_ms = stats[0]
        per = dict(zip(phases, stats[2:]))
        hot = per["chamfer_fwd"] + per["emd_fwd"]
        report = {
            "ops": ops, "n_gpus": world, "batch_per_gpu": batch, "image": image, "points": 1024, "steps": steps,
            "params_M": round(nparam / 1e6, 1), "ms_per_step": wall_ms, "samples_per_s": world * batch / (wall_ms * 1e-3),
            "phase_ms": {k: round(v, 3) for k, v in per.items()},
            "hot_path_fwd_share": hot / max(sum(per.values()), 1e-9),
            "loss_first_step": first_loss, "loss": loss_v,
            "grad_allreduce": None if world == 1 else {
                "bytes": 4 * nparam, "step_ms_without_allreduce": stats[1], "exposed_ms": wall_ms - stats[1],
                "note": "DDP gradient all-reduce over NCCL/NVLink; exposed = step time minus the same step under no_sync()"},
            "what": "generator (plain torch, random init) fwd -> 100*CD + 100*EMD(eps 0.05, 3000 iters) [+ projection terms, "
                    "no gradient] -> backward -> Adam; max over ranks; backward includes the ops' own backward kernels",
        }
        if not quiet:
            print(json.dumps(report))
    if own_pg:
        dist.destroy_process_group()
    return report


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="ours", choices=["ours", "reference"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--no-proj", action="store_true")
    a = ap.parse_args()
    run(a.ops, a.steps, a.warmup, a.batch, a.image, not a.no_proj)
