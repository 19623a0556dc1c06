"""The boundary on both sides (VERDICT r1 missing 5): pybind/psd_pybind.cpp compiled into the reference's two native module
names (`chamfer_3D`, `emd`) over the C ABI, and the reference's UNMODIFIED metric/chamfer3D/dist_chamfer_3D.py and
metric/emd/emd_module.py (verbatim copies in the git-ignored oracle/_ref/py, made by oracle/build_ref.py) running on top of it,
compared with the reference's own extensions (oracle/_ref) and the CPU oracle."""
import importlib
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import make_clouds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "3d-pointcloudreconstruction_b200", "pybind", "_build")
REFPY = os.path.join(ROOT, "oracle", "_ref", "py")


def test_pybind_modules_are_built_and_export_the_reference_names():
    """CPU check: both modules exist (built by __graft_entry__.build()) and export forward / backward.  Imported in a
    subprocess so that the module name `emd` / `chamfer_3D` cannot collide with anything loaded in this session."""
    for name in ("chamfer_3D", "emd"):
        assert os.path.exists(os.path.join(BUILD, name + ".so")), f"{name}.so missing: run python 3d-pointcloudreconstruction_b200/build.py --pybind"
    code = ("import sys, torch; sys.path.insert(0, %r); import chamfer_3D, emd; "
            "assert callable(chamfer_3D.forward) and callable(chamfer_3D.backward) and callable(emd.forward) and callable(emd.backward); "
            "print(chamfer_3D.__file__)" % BUILD)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert BUILD in out.stdout


_WORKER = r'''
import importlib, importlib.util, os, sys
import numpy as np, torch
ROOT, BUILD, REFPY = sys.argv[1:4]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
importlib.find_loader = lambda name: importlib.util.find_spec(name)      # py3.12 shim for dist_chamfer_3D.py:6 (INTEGRATION.md 2)
sys.path.insert(0, BUILD)                                                # native modules: THIS repo's pybind build
sys.path.insert(0, REFPY)                                                # wrappers: the reference's files, unmodified
import chamfer_3D, emd
assert os.path.dirname(chamfer_3D.__file__) == BUILD and os.path.dirname(emd.__file__) == BUILD
import dist_chamfer_3D, emd_module
assert os.path.dirname(dist_chamfer_3D.__file__) == REFPY and os.path.dirname(emd_module.__file__) == REFPY
from conftest import make_clouds
from oracle import oracle as O
O.build_c()
dev = torch.device("cuda:0")
# ---- chamfer: reference wrapper over our module == oracle, forward bits and gradients
x, y = make_clouds("uniform", 8, 1024, 1500, seed=3)
tx = torch.from_numpy(x).to(dev).requires_grad_(True); ty = torch.from_numpy(y).to(dev).requires_grad_(True)
d1, d2, i1, i2 = dist_chamfer_3D.chamfer_3DDist()(tx, ty)
(d1.mean() + d2.mean()).backward()
torch.cuda.synchronize()
w = O.chamfer_forward(x, y, nthreads=8)
for got, want in zip((d1, d2, i1, i2), w):
    assert np.array_equal(got.detach().cpu().numpy(), want), "chamfer forward"
g1 = np.full(d1.shape, 1.0 / d1.numel(), np.float32); g2 = np.full(d2.shape, 1.0 / d2.numel(), np.float32)
wg1, wg2 = O.chamfer_backward(x, y, g1, g2, w[2], w[3])
assert np.allclose(tx.grad.cpu().numpy(), wg1, rtol=1e-5, atol=1e-9) and np.allclose(ty.grad.cpu().numpy(), wg2, rtol=1e-5, atol=1e-9)
# ---- the same against the reference's own extension (oracle/_ref), raw module level
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_chamfer_3D")); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ref_emd"))
import ref_chamfer_3D, ref_emd
z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
rd1, rd2, ri1, ri2 = z(8, 1024), z(8, 1500), z(8, 1024, dt=torch.int32), z(8, 1500, dt=torch.int32)
ref_chamfer_3D.forward(tx.detach(), ty.detach(), rd1, rd2, ri1, ri2)
torch.cuda.synchronize()
assert torch.equal(rd1, d1.detach()) and torch.equal(rd2, d2.detach()) and torch.equal(ri1, i1) and torch.equal(ri2, i2)
# ---- EMD: reference wrapper (its 12 scratch tensors and all) over our module == oracle
a, b_ = make_clouds("uniform", 4, 1024, 1024, seed=5)
ta = torch.from_numpy(a).to(dev).requires_grad_(True); tb = torch.from_numpy(b_).to(dev)
dist, ass = emd_module.emdModule()(ta, tb, 0.005, 50)
torch.sqrt(dist).mean(1).mean().backward()
torch.cuda.synchronize()
wd, wa = O.emd_forward(a, b_, 0.005, 50, nthreads=4)[:2]
assert np.array_equal(ass.cpu().numpy(), wa) and np.array_equal(dist.detach().cpu().numpy(), wd)
gd = (1.0 / 4 / 1024) / (2 * np.sqrt(wd))
assert np.allclose(ta.grad.cpu().numpy(), O.emd_backward(a, b_, gd.astype(np.float32), wa), rtol=1e-5, atol=1e-9)
# the shape checks of the native module are the reference's (emd_cuda.cu:236-249): -1, nothing raised at this level
bad = torch.rand(1, 1000, 3, device=dev)
s = [z(1, 1000), z(1, 1000, dt=torch.int32), z(1, 1000), z(1, 1000, dt=torch.int32), z(1, 1000, dt=torch.int32), z(1, 1000), z(1, 1000),
     z(1000, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(1000, dt=torch.int32)]
assert emd.forward(bad, bad, *s, 0.005, 5) == -1
print("PYBIND_OK")
'''


@pytest.mark.gpu
def test_reference_wrappers_run_unmodified_on_the_pybind_modules(cuda):
    if not os.path.isdir(REFPY):
        pytest.skip("oracle/_ref/py missing (oracle/build_ref.py needs /root/reference)")
    out = subprocess.run([sys.executable, "-c", _WORKER, ROOT, BUILD, REFPY], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "PYBIND_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
