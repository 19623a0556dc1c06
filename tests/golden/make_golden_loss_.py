"""Golden vectors for the loss_.py mirror, produced by the REFERENCE's own loss/loss_.py run on the CPU in the build container
(geomloss and the CUDA wrapper are stubbed: neither is used by the functions called here).

    python tests/golden/make_golden_loss_.py      # needs /root/reference; writes tests/golden/loss__*_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from conftest import make_clouds  # noqa: E402

for name, attrs in (("geomloss", {"SamplesLoss": object}), ("dist_chamfer_3D", {"chamfer_3DDist": object})):
    stub = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(stub, k, v)
    sys.modules[name] = stub
spec = importlib.util.spec_from_file_location("ref_loss_", "/root/reference/loss/loss_.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

for kind, (b, n, m), thr in (("uniform", (3, 700, 900), 1e-3), ("clustered", (2, 1024, 1024), 1e-4), ("dup", (2, 513, 640), 1e-4),
                             ("lattice", (2, 300, 200), 0.02)):
    x, y = make_clouds(kind, b, n, m, seed=123)
    tx, ty = torch.from_numpy(x), torch.from_numpy(y)
    loss, mins1, mins2 = ref.batch_NN_loss(tx, ty)
    d1, d2, i1, i2 = ref.distChamfer(tx, ty)
    f, p1, p2 = ref.fscore(tx, ty, thr)
    out = os.path.join(HERE, f"loss__{kind}_{b}x{n}x{m}_ref.npz")
    np.savez_compressed(out, x=x, y=y, thr=np.float64(thr), nn_loss=loss.numpy(), mins1=mins1.numpy(), mins2=mins2.numpy(),
                        d1=d1.numpy(), d2=d2.numpy(), i1=i1.numpy(), i2=i2.numpy(), fscore=f.detach().numpy(), p1=p1.numpy(), p2=p2.numpy())
    print(out, float(loss), float(f), float(p1), float(p2))
