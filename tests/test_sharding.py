"""CPU tests (gloo, world_size 2) of the multi-GPU decomposition: the partitioning and reduction logic of
`sharding.py` with the CPU oracle injected as the local operator must reproduce the single-process result."""
import os
import sys
from importlib import import_module

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleOps:
    """Local operator for the CPU tests: the oracle restricted to a query slice (same contract as
    psd_chamfer_forward_ex: outputs indexed by global query index, zeros outside the slice)."""

    def forward_slice(self, xyz1, xyz2, q_begin, q_count, fs_thr):
        from oracle import oracle as O
        x, y = xyz1.numpy(), xyz2.numpy()
        n, m = x.shape[1], y.shape[1]
        d1, d2, i1, i2 = O.chamfer_forward(x, y)
        out, cnts, sums = {}, [], []
        for name, dd_full, ii_full, size in (("1", d1, i1, n), ("2", d2, i2, m)):
            lo, hi = min(q_begin, size), min(q_begin + q_count, size)
            dd = np.zeros_like(dd_full)
            ii = np.zeros_like(ii_full)
            dd[:, lo:hi] = dd_full[:, lo:hi]
            ii[:, lo:hi] = ii_full[:, lo:hi]
            out["dist" + name], out["idx" + name] = torch.from_numpy(dd), torch.from_numpy(ii)
            cnts.append((dd_full[:, lo:hi] < fs_thr).sum(1))
            sums.append(dd.astype(np.float64).sum(1))
        out["sums"] = torch.from_numpy(np.stack(sums, 1).astype(np.float32))
        out["counts"] = torch.from_numpy(np.stack(cnts, 1).astype(np.int32))
        return out

    def backward(self, xyz1, xyz2, graddist1, graddist2, idx1, idx2):
        from oracle import oracle as O
        g1, g2 = O.chamfer_backward(xyz1.numpy(), xyz2.numpy(), graddist1.numpy(), graddist2.numpy(), idx1.numpy(), idx2.numpy())
        return torch.from_numpy(g1), torch.from_numpy(g2)


def _sharding_module():
    import psd_b200
    psd_b200.load()
    return import_module(psd_b200.PKG_NAME + ".sharding")


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import make_clouds
        from oracle import oracle as O
        sh = _sharding_module()
        ok = {}
        # ---- query sharding, ragged sizes (n != m, not divisible by world)
        x, y = make_clouds("clustered", 3, 301, 517, seed=5)
        tx, ty = torch.from_numpy(x), torch.from_numpy(y)
        out = sh.chamfer_query_sharded(tx, ty, rank, world, threshold=1e-3, ops=OracleOps())
        d1, d2, i1, i2 = O.chamfer_forward(x, y)
        ok["assemble"] = all(np.array_equal(out[k].numpy(), w) for k, w in (("dist1", d1), ("dist2", d2), ("idx1", i1), ("idx2", i2)))
        out_ar = sh.chamfer_query_sharded(tx, ty, rank, world, threshold=1e-3, ops=OracleOps(), assemble="all_reduce")
        ok["assemble_all_reduce"] = all(np.array_equal(out_ar[k].numpy(), w) for k, w in (("dist1", d1), ("dist2", d2), ("idx1", i1), ("idx2", i2)))
        # very asymmetric clouds: the shorter direction's queries all belong to rank 0 (rank 1 owns an empty slice of it)
        xa, ya = make_clouds("uniform", 2, 700, 40, seed=6)
        oa = sh.chamfer_query_sharded(torch.from_numpy(xa), torch.from_numpy(ya), rank, world, ops=OracleOps())
        wa = O.chamfer_forward(xa, ya)
        ok["asymmetric"] = all(np.array_equal(oa[k].numpy(), w) for k, w in zip(("dist1", "dist2", "idx1", "idx2"), wa))
        c1, c2 = O.fscore_counts(d1, d2, 1e-3)
        ok["counts"] = np.array_equal(out["counts"].numpy(), np.stack([c1, c2], 1))
        ok["sums"] = np.allclose(out["sums"].numpy(), np.stack([d1.sum(1), d2.sum(1)], 1), rtol=1e-5)
        f, _, _ = O.fscore_from_counts(c1, c2, 301, 517)
        ok["fscore"] = abs(float(out["fscore"].mean()) - float(f)) < 1e-6
        # ---- gradient under query sharding == single-process gradient
        g1 = np.random.default_rng(1).random((3, 301), dtype=np.float32)
        g2 = np.random.default_rng(2).random((3, 517), dtype=np.float32)
        gx, gy = sh.chamfer_backward_query_sharded(tx, ty, torch.from_numpy(g1), torch.from_numpy(g2), out["idx1"], out["idx2"],
                                                   rank, world, ops=OracleOps())
        wx, wy = O.chamfer_backward(x, y, g1, g2, i1, i2)
        ok["grad"] = np.allclose(gx.numpy(), wx, rtol=1e-5, atol=1e-6) and np.allclose(gy.numpy(), wy, rtol=1e-5, atol=1e-6)
        # ---- batch sharding: rank-local clouds, loss all-reduce
        xb, yb = make_clouds("uniform", 5, 128, 160, seed=9)
        D1, D2, _, _ = O.chamfer_forward(xb, yb)
        lx = sh.batch_shard(torch.from_numpy(xb), world, rank).numpy()
        ly = sh.batch_shard(torch.from_numpy(yb), world, rank).numpy()
        l1, l2, _, _ = O.chamfer_forward(lx, ly)
        _, glob = sh.chamfer_loss_batch_sharded(torch.from_numpy(l1), torch.from_numpy(l2), 5)
        ok["batch_loss"] = abs(float(glob) - float(D1.mean() + D2.mean())) < 1e-6
        xe, ye = make_clouds("uniform", 3, 1024, 1024, seed=11)
        de = O.emd_forward(xe, ye, 0.05, 20)[0]
        le = O.emd_forward(sh.batch_shard(torch.from_numpy(xe), world, rank).numpy(),
                           sh.batch_shard(torch.from_numpy(ye), world, rank).numpy(), 0.05, 20)[0]
        _, ge = sh.emd_loss_batch_sharded(torch.from_numpy(le), 3)
        ok["emd_loss"] = abs(float(ge) - float(np.sqrt(de).mean(1).mean())) < 1e-6
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_split_range_partitions():
    sh = _sharding_module()
    for total in (0, 1, 7, 32, 131072, 1000):
        for world in (1, 2, 3, 4, 8):
            parts = [sh.split_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_world_size_2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        for k, v in ret[rank].items():
            assert v, f"rank {rank}: {k} failed"
