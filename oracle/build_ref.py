"""Compile the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

The sources are compiled where they lie under /root/reference (nothing enters this repo's history):
    metric/chamfer3D/{chamfer_cuda.cpp,chamfer3D.cu}  -> oracle/_ref/ref_chamfer_3D.so
    metric/emd/{emd.cpp,emd_cuda.cu}                  -> oracle/_ref/ref_emd.so
    metric/chamfer3D/dist_chamfer_3D.py, metric/emd/emd_module.py, loss/loss_.py -> oracle/_ref/py/ (verbatim copies, so that
    the GPU box -- which has no /root/reference -- can run the reference's own Python on top of either native module)
Flags are the ones torch's BuildExtension would pass for the reference's own setup.py (no
--use_fast_math: -fmad=true, IEEE sqrt/div, no FTZ), with the arch pinned to sm_100a.  The module
names carry a ref_ prefix (TORCH_EXTENSION_NAME) so they can be imported next to this repo's drop-in
`chamfer_3D` / `emd` modules.  oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.

Only tests/, __graft_entry__.py and bench.py's reference/cpu_baseline legs may load these modules.
Run: python oracle/build_ref.py     (needs /root/reference; a no-op with a notice when it is absent)
"""
import os
import sys

REF = os.environ.get("PSD_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(verbose: bool = False) -> bool:
    if not os.path.isdir(os.path.join(REF, "metric")):
        print(f"[build_ref] {REF} not present: keeping whatever is already in {OUT}")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    cuda_flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
    for name, sub, files in (
        ("ref_chamfer_3D", "metric/chamfer3D", ["chamfer_cuda.cpp", "chamfer3D.cu"]),
        ("ref_emd", "metric/emd", ["emd.cpp", "emd_cuda.cu"]),
    ):
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(
            name=name,
            sources=[os.path.join(REF, sub, f) for f in files],
            build_directory=bdir,
            extra_cuda_cflags=cuda_flags,
            extra_cflags=["-O2"],
            is_python_module=True,
            verbose=verbose,
        )
        print(f"[build_ref] built {name} in {bdir}")
    # the reference's own Python wrappers / CPU chamfer, byte for byte, next to its compiled extensions (git-ignored like
    # them): tests run them UNMODIFIED on top of this repo's pybind modules, bench.py --impl reference times loss_.py
    import shutil
    pdir = os.path.join(OUT, "py")
    os.makedirs(pdir, exist_ok=True)
    for rel in ("metric/chamfer3D/dist_chamfer_3D.py", "metric/emd/emd_module.py", "loss/loss_.py"):
        shutil.copy2(os.path.join(REF, rel), os.path.join(pdir, os.path.basename(rel)))
    print(f"[build_ref] copied the reference's Python wrappers to {pdir}")
    return True


if __name__ == "__main__":
    build(verbose="-v" in sys.argv)
