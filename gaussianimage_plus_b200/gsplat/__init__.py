"""Drop-in replacement for the reference's `gsplat` package, hot path only.

Same import surface as gsplat/gsplat/__init__.py:3-17 for the 2-D path:

    from gsplat.project_gaussians_2d_covariance import project_gaussians_2d_covariance
    from gsplat.rasterize_sum_plus import rasterize_gaussians_plus
    from gsplat import bin_and_sort_gaussians, compute_cumulative_intersects, ...

Call `gaussianimage_plus_b200.install_as_gsplat()` to register this package under the name
`gsplat`.  The 3-D pipeline (project_gaussians, rasterize_gaussians, sh) is
out of scope and raises on use.
"""
from .project_gaussians_2d import project_gaussians_2d
from .project_gaussians_2d_covariance import project_gaussians_2d_covariance
from .project_gaussians_2d_scale_rot import project_gaussians_2d_scale_rot
from .rasterize_sum import rasterize_gaussians_sum
from .rasterize_sum_plus import rasterize_gaussians_plus
from .utils import (bin_and_sort_gaussians, compute_cov2d_bounds, compute_cumulative_intersects,
                    get_tile_bin_edges, map_gaussian_to_intersects)

__version__ = "1.1.3+gi2d"


def _out_of_scope(name):
    def fn(*a, **k):
        raise NotImplementedError(f"gsplat.{name} belongs to the 3-D pipeline, which is outside the hot path "
                                  "this package replaces (SURVEY.md 2.2 #14)")
    fn.__name__ = name
    return fn


project_gaussians = _out_of_scope("project_gaussians")
rasterize_gaussians = _out_of_scope("rasterize_gaussians")
spherical_harmonics = _out_of_scope("spherical_harmonics")

__all__ = [
    "__version__", "project_gaussians", "project_gaussians_2d", "project_gaussians_2d_scale_rot",
    "project_gaussians_2d_covariance", "rasterize_gaussians", "rasterize_gaussians_sum",
    "rasterize_gaussians_plus", "spherical_harmonics", "bin_and_sort_gaussians",
    "compute_cumulative_intersects", "compute_cov2d_bounds", "get_tile_bin_edges", "map_gaussian_to_intersects",
]
