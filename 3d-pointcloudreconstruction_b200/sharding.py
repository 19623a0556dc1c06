"""Multi-GPU decomposition of the hot path (one process per GPU, torch.distributed; NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The reference has no distributed layer at all, so this
file defines the only exchange steps the path needs:

* batch sharding (default): rank r owns clouds [start, start+count) end to end -- forward, backward, EMD.
  No data-path collective; one all_reduce of a tiny vector for the reported loss / F-score.
* query sharding (large clouds, B < world): every rank holds the full clouds, runs the NN search for a 1/world
  slice of the QUERY points of both directions (psd_chamfer_forward_ex q_begin/q_count), then one all_reduce(SUM) each
  for the per-cloud sums and the F-score counts and, when the full dist/idx are wanted on every rank, an
  all_gather of the owned slices (or an all_reduce of zero-padded full tensors).  Backward: each rank scatters the gradient terms of its own query slice, all_reduce(SUM) of grad_xyz.
  EMD is never query-sharded (the auction state is per cloud and iterative).

The local operator is injectable (`ops`) so that the partitioning / reduction logic can be tested on CPU
against a single-process run; the default is the CUDA implementation of this package."""
import torch
import torch.distributed as dist


def split_range(total: int, world: int, rank: int):
    """Contiguous, balanced split of range(total): returns (start, count)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def batch_shard(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """Rank-local clouds of a [B, ...] tensor."""
    s, c = split_range(t.shape[0], world, rank)
    return t[s:s + c]


def _all_reduce(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def chamfer_loss_batch_sharded(dist1_local, dist2_local, b_global: int, group=None):
    """Global `mean(dist1) + mean(dist2)` (loss/loss.py:36) from rank-local distances.

    Returns (loss_for_backward, loss_global): the first is the rank-local share
    sum(dist_local) / (B_global * N) -- its gradient w.r.t. the local clouds IS the gradient of the global
    loss, no gradient collective needed for the op itself; the second is the all-reduced value for logging."""
    n, m = dist1_local.shape[1], dist2_local.shape[1]
    local = dist1_local.sum() / (b_global * n) + dist2_local.sum() / (b_global * m)
    glob = _all_reduce(local.detach().clone().reshape(1), group)[0]
    return local, glob


def emd_loss_batch_sharded(dist_local, b_global: int, group=None):
    """Global `sqrt(dist).mean(1).mean()` (loss/loss.py:25) from the rank-local EMD distances."""
    n = dist_local.shape[1]
    local = torch.sqrt(dist_local).sum() / (b_global * n)
    glob = _all_reduce(local.detach().clone().reshape(1), group)[0]
    return local, glob


class _CudaOps:
    """Default local operator: this package's CUDA kernels."""

    def forward_slice(self, xyz1, xyz2, q_begin, q_count, fs_thr):
        try:
            from . import chamfer_3D, _lib
        except ImportError:
            import chamfer_3D
            import _lib
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        out = {
            "dist1": torch.zeros(b, n, device=dev), "dist2": torch.zeros(b, m, device=dev),
            "idx1": torch.zeros(b, n, device=dev, dtype=torch.int32), "idx2": torch.zeros(b, m, device=dev, dtype=torch.int32),
            "sums": torch.zeros(b, 2, device=dev), "counts": torch.zeros(b, 2, device=dev, dtype=torch.int32),
        }
        rc = chamfer_3D.forward_ex(xyz1.contiguous(), xyz2.contiguous(), out["dist1"], out["dist2"], out["idx1"], out["idx2"],
                                   sums=out["sums"], fs_thr=fs_thr, fs_count=out["counts"], q_begin=q_begin, q_count=q_count)
        _lib.raise_on_cuda_error(rc, "chamfer_3D.forward_ex")
        return out

    def backward(self, xyz1, xyz2, graddist1, graddist2, idx1, idx2):
        try:
            from . import chamfer_3D, _lib
        except ImportError:
            import chamfer_3D
            import _lib
        g1 = torch.zeros_like(xyz1)
        g2 = torch.zeros_like(xyz2)
        rc = chamfer_3D.backward(xyz1.contiguous(), xyz2.contiguous(), g1, g2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)
        _lib.raise_on_cuda_error(rc, "chamfer_3D.backward")
        return g1, g2


def chamfer_query_sharded(xyz1, xyz2, rank: int, world: int, threshold: float = 1e-4, group=None, ops=None,
                          assemble=True):
    """Query-sharded chamfer + F-score (BASELINE.json configs[4]).  Every rank passes the SAME full clouds.

    Returns a dict with the all-reduced per-cloud `sums` [B,2] and `counts` [B,2], the derived `chamfer` [B],
    `fscore` [B], `precision_1/2` [B], the rank's query slices and, if `assemble`, the full `dist1/dist2/idx1/idx2` on
    every rank: assemble=True / "all_gather" gathers the owned slices (one collective per tensor, 1/world of the bytes),
    assemble="all_reduce" sums full-length tensors that are zero outside the owned slice.  Each direction's queries are
    split on their own (n and m may differ)."""
    ops = ops or _CudaOps()
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    # one (begin, count) pair addresses both directions in psd_chamfer_forward_ex: split the longer cloud; the shorter
    # direction is clipped to its own length by the kernel launcher
    q_begin, q_count = split_range(max(n, m), world, rank)
    out = ops.forward_slice(xyz1, xyz2, q_begin, q_count, threshold)
    sums = _all_reduce(out["sums"], group)
    counts = _all_reduce(out["counts"], group)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if assemble and multi:
        if assemble == "all_reduce":
            for k in ("dist1", "dist2", "idx1", "idx2"):
                _all_reduce(out[k], group)
        else:
            for k, total in (("dist1", n), ("dist2", m), ("idx1", n), ("idx2", m)):
                # ranks split max(n, m); the shorter cloud's slices are the same ranges clipped to its length
                per = -(-max(n, m) // world)
                b_ = out[k].shape[0]
                send = torch.zeros(b_, per, device=out[k].device, dtype=out[k].dtype)
                lo, hi = min(q_begin, total), min(q_begin + q_count, total)
                if hi > lo:
                    send[:, : hi - lo] = out[k][:, lo:hi]
                recv = torch.empty(world * b_, per, device=out[k].device, dtype=out[k].dtype)   # rank-major concatenation
                dist.all_gather_into_tensor(recv, send, group=group)
                recv = recv.view(world, b_, per)
                full = torch.empty_like(out[k])
                for r in range(world):
                    rb, rc = split_range(max(n, m), world, r)
                    lo_r, hi_r = min(rb, total), min(rb + rc, total)
                    if hi_r > lo_r:
                        full[:, lo_r:hi_r] = recv[r, :, : hi_r - lo_r]
                out[k] = full
    p1 = counts[:, 0].float() / n
    p2 = counts[:, 1].float() / m
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    out.update({"sums": sums, "counts": counts, "chamfer": sums[:, 0] / n + sums[:, 1] / m, "fscore": f,
                "precision_1": p1, "precision_2": p2, "q_begin": q_begin, "q_count": q_count})
    return out


def chamfer_backward_query_sharded(xyz1, xyz2, graddist1, graddist2, idx1, idx2, rank: int, world: int, group=None,
                                   ops=None):
    """Gradient under query sharding: the rank keeps only the upstream gradients of ITS query slice (zeros
    elsewhere contribute nothing), scatters them with the normal backward kernel, and the per-rank partial
    gradients are summed with one all_reduce of [B,N,3] + [B,M,3]."""
    ops = ops or _CudaOps()
    n, m = xyz1.shape[1], xyz2.shape[1]
    q_begin, q_count = split_range(max(n, m), world, rank)
    mask1 = torch.zeros_like(graddist1)
    mask2 = torch.zeros_like(graddist2)
    mask1[:, q_begin:min(q_begin + q_count, n)] = 1
    mask2[:, q_begin:min(q_begin + q_count, m)] = 1
    g1, g2 = ops.backward(xyz1, xyz2, graddist1 * mask1, graddist2 * mask2, idx1, idx2)
    _all_reduce(g1, group)
    _all_reduce(g2, group)
    return g1, g2
