"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference gsplat CUDA extension.

The reference (Sweethyh/GaussianImage_plus) ships its hot path as a pybind11/torch
CUDA extension (`gsplat/gsplat/cuda/csrc/*.cu,*.cpp`, loaded by
`gsplat/gsplat/cuda/_backend.py:54-94`).  This recipe compiles those sources *where
they lie* under /root/reference (nothing is copied into the repo) for sm_100a and
drops the resulting shared objects into `oracle/_ref/<variant>/` (git-ignored, but
shipped to the GPU box by gpurun).  Two variants:

  o3        -O3                      -- flags of the JIT path  (`_backend.py:39-40`)
  fastmath  -O3 --use_fast_math      -- flags of `pip install` (`gsplat/setup.py:79`)

The GPU parity tests (`tests/test_ref_cuda_parity.py`) and `bench.py`'s `ref_cuda`
leg load them through `oracle/ref_cuda.py`.  Nothing in the product package may
import this module.

Usage:  python oracle/build_ref.py [o3] [fastmath]
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/gsplat/gsplat/cuda/csrc"
VARIANTS = {
    "o3": ["-O3"],
    "fastmath": ["-O3", "--use_fast_math", "--expt-relaxed-constexpr"],
}


def so_path(variant: str) -> str:
    return os.path.join(HERE, "_ref", variant, f"gsplat_ref_{variant}.so")


def build(variant: str, verbose: bool = True) -> str:
    """Compile the reference extension for sm_100a; returns the .so path."""
    out = so_path(variant)
    if os.path.exists(out):
        return out
    if not os.path.isdir(REF_CSRC):
        raise FileNotFoundError(f"{REF_CSRC} not present (GPU box?) and {out} was not prebuilt")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    bdir = os.path.dirname(out)
    os.makedirs(bdir, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(REF_CSRC, "*.cu"))) + sorted(
        glob.glob(os.path.join(REF_CSRC, "*.cpp"))
    )
    load(
        name=f"gsplat_ref_{variant}",
        sources=sources,
        extra_cflags=["-O3"],
        extra_cuda_cflags=VARIANTS[variant],
        extra_include_paths=[os.path.join(REF_CSRC, "third_party", "glm")],
        build_directory=bdir,
        verbose=verbose,
        is_python_module=False,
    )
    # drop the objects, keep the .so
    for f in glob.glob(os.path.join(bdir, "*.o")):
        os.remove(f)
    return out


if __name__ == "__main__":
    for v in sys.argv[1:] or ["o3", "fastmath"]:
        print(build(v))
