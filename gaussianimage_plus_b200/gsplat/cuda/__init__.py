"""`gsplat.cuda` of the drop-in package: the reference resolves `_C.<name>` lazily against its
pybind11 extension (`gsplat/gsplat/cuda/__init__.py:4-11`); here the same names resolve against
the ctypes binding of libgi2d (gaussianimage_plus_b200/binding.py).  No JIT, no fallback."""
from ... import binding as _binding

_NAMES = [
    "project_gaussians_2d_forward", "project_gaussians_2d_backward",
    "project_gaussians_2d_scale_rot_forward", "project_gaussians_2d_scale_rot_backward",
    "project_gaussians_2d_covariance_forward", "project_gaussians_2d_covariance_backward",
    "compute_cov2d_bounds", "map_gaussian_to_intersects", "get_tile_bin_edges",
    "rasterize_sum_forward", "rasterize_sum_backward",
    "rasterize_sum_plus_forward", "rasterize_sum_plus_backward",
]
for _n in _NAMES:
    globals()[_n] = getattr(_binding, _n)
__all__ = list(_NAMES)
