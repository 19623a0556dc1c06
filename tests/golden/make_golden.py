"""Generate the committed golden vectors.

    python tests/golden/make_golden.py            # CPU: outputs from the oracle   (source = "oracle")
    python tests/golden/make_golden.py --ref      # GPU box: outputs from the reference's own CUDA extensions
                                                  # built unmodified into oracle/_ref (source = "reference_cuda");
                                                  # written to gpurun_out/golden/ and copied into tests/golden/.
Inputs are small seeded clouds (incl. ties / duplicates / ragged sizes); files stay a few hundred KB."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_clouds  # noqa: E402

CHAMFER = [("uniform", 2, 600, 700, 1), ("lattice", 2, 300, 520, 2), ("dup", 1, 513, 1030, 3), ("clustered", 2, 1000, 257, 4)]
EMD = [("uniform", 2, 1024, 0.005, 50, 5), ("clustered", 1, 1024, 0.05, 200, 6), ("uniform", 1, 2048, 0.005, 50, 7)]


def main():
    use_ref = "--ref" in sys.argv
    out_dir = os.path.join(ROOT, "gpurun_out", "golden") if use_ref else HERE
    os.makedirs(out_dir, exist_ok=True)
    if use_ref:
        import torch
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_chamfer_3D"))
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_emd"))
        import ref_chamfer_3D, ref_emd
        dev = torch.device("cuda:0")
    from oracle import oracle as O
    tag = "ref" if use_ref else "oracle"
    for kind, b, n, m, seed in CHAMFER:
        x, y = make_clouds(kind, b, n, m, seed)
        if use_ref:
            tx, ty = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
            d1 = torch.zeros(b, n, device=dev); d2 = torch.zeros(b, m, device=dev)
            i1 = torch.zeros(b, n, device=dev, dtype=torch.int32); i2 = torch.zeros(b, m, device=dev, dtype=torch.int32)
            ref_chamfer_3D.forward(tx, ty, d1, d2, i1, i2)
            torch.cuda.synchronize()
            outs = [t.cpu().numpy() for t in (d1, d2, i1, i2)]
        else:
            outs = O.chamfer_forward(x, y)
        meta = {"op": "chamfer", "kind": kind, "seed": seed, "source": "reference_cuda" if use_ref else "oracle"}
        np.savez_compressed(os.path.join(out_dir, f"chamfer_{kind}_{b}x{n}x{m}_{tag}.npz"), xyz1=x, xyz2=y, dist1=outs[0],
                            dist2=outs[1], idx1=outs[2], idx2=outs[3], meta=json.dumps(meta))
    for kind, b, n, eps, iters, seed in EMD:
        x, y = make_clouds(kind, b, n, n, seed)
        od, oa, st = O.emd_forward(x, y, eps, iters, want_stats=True)
        exact = True
        if use_ref:
            tx, ty = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
            z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
            dist = z(b, n); ass = z(b, n, dt=torch.int32) - 1; inv = z(b, n, dt=torch.int32) - 1; price = z(b, n)
            bid = z(b, n, dt=torch.int32); bi = z(b, n); mi = z(b, n); ui = z(b * n, dt=torch.int32); mx = z(b * n, dt=torch.int32)
            c1 = z(512, dt=torch.int32); c2 = z(512, dt=torch.int32); c3 = z(512, dt=torch.int32)
            ref_emd.forward(tx, ty, dist, ass, price, inv, bid, bi, mi, ui, c1, c2, c3, mx, eps, iters)
            torch.cuda.synchronize()
            d, a = dist.cpu().numpy(), ass.cpu().numpy()
            exact = bool(np.array_equal(a, oa))  # False only if the reference's GetMax race picked another winner
            print(f"emd {kind} n={n}: reference == oracle: {exact}; oracle multi-winner events: {st['multi_winner']}")
        else:
            d, a = od, oa
        meta = {"op": "emd", "kind": kind, "seed": seed, "eps": eps, "iters": iters, "exact": exact,
                "multi_winner_events": st["multi_winner"], "source": "reference_cuda" if use_ref else "oracle"}
        np.savez_compressed(os.path.join(out_dir, f"emd_{kind}_{b}x{n}_{tag}.npz"), xyz1=x, xyz2=y, dist=d, assignment=a,
                            meta=json.dumps(meta))
    print("wrote golden vectors to", out_dir)


if __name__ == "__main__":
    main()
