"""gaussianimage_plus_b200 -- a B200 (sm_100a) native 2-D Gaussian image rasterizer.

The hot path of Sweethyh/GaussianImage_plus (project -> bin/sort -> rasterize-sum forward ->
backward -> project backward + Adam) as hand-written CUDA behind a C ABI (include/gi2d.h,
libgi2d.so), exposed to PyTorch under the reference's own operator API:

  gaussianimage_plus_b200.gsplat      drop-in for the reference's `gsplat` package (2-D path)
  gaussianimage_plus_b200.binding     the `_C` replacement (tensor-level binding of the C ABI)
  gaussianimage_plus_b200.fit         fused, graph-captured fit loop (GaussianImageFitter)
  gaussianimage_plus_b200.parallel    image-set sharding and tile-row split across GPUs

There is no CPU fallback: importing is cheap, the first operator call loads libgi2d.so and
raises if it is missing.
"""
import sys as _sys

__version__ = "0.1.0"


def install_as_gsplat() -> None:
    """Register the drop-in package under the name `gsplat`, so that the reference's models
    (`from gsplat.project_gaussians_2d_covariance import ...`, models/gaussianimage_covariance.py:2-3)
    run unmodified on the new kernels.  Call before importing the reference's `models`."""
    import importlib

    pkg = importlib.import_module(__name__ + ".gsplat")
    _sys.modules["gsplat"] = pkg
    for sub in ("project_gaussians_2d", "project_gaussians_2d_covariance", "project_gaussians_2d_scale_rot",
                "rasterize_sum", "rasterize_sum_plus", "utils", "cuda"):
        _sys.modules["gsplat." + sub] = importlib.import_module(f"{__name__}.gsplat.{sub}")
