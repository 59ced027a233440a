"""CPU, only where /root/reference is mounted (this container; the GPU box skips): the drop-in `gsplat` package has
the reference's public signatures -- every function the models call, parameter names, order and defaults compared
with `inspect.signature` -- and the reference's OWN model file imports on top of `install_as_gsplat()` (the four
third-party packages absent here are stubbed, as in tests/golden/make_golden_quant.py) and finds every name it
needs.  Nothing is computed: no GPU, no oracle."""
import importlib
import importlib.util
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gsplat", "gsplat")),
                                reason="reference sources not mounted")

FUNCS = {
    "project_gaussians_2d_covariance": ["project_gaussians_2d_covariance"],
    "project_gaussians_2d": ["project_gaussians_2d"],
    "project_gaussians_2d_scale_rot": ["project_gaussians_2d_scale_rot"],
    "rasterize_sum_plus": ["rasterize_gaussians_plus"],
    "rasterize_sum": ["rasterize_gaussians_sum"],
    "utils": ["bin_and_sort_gaussians", "compute_cumulative_intersects", "map_gaussian_to_intersects",
              "get_tile_bin_edges", "compute_cov2d_bounds"],
}


def _load_reference_module(name):
    """One module of the reference's gsplat package, loaded by path under a private package name (its relative
    `import gsplat.cuda as _C` resolves lazily, so importing needs no compiled extension)."""
    pkg_name = "_ref_gsplat"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REF, "gsplat", "gsplat")]
        sys.modules[pkg_name] = pkg
        cuda = types.ModuleType(pkg_name + ".cuda")       # `import gsplat.cuda as _C`: only attribute access at call time
        sys.modules[pkg_name + ".cuda"] = cuda
        pkg.cuda = cuda
    src = open(os.path.join(REF, "gsplat", "gsplat", name + ".py")).read()
    src = src.replace("import gsplat.cuda as _C", f"import {pkg_name}.cuda as _C").replace(
        "from .utils import", f"from {pkg_name}.utils import")
    mod = types.ModuleType(f"{pkg_name}.{name}")
    mod.__package__ = pkg_name
    sys.modules[f"{pkg_name}.{name}"] = mod
    exec(compile(src, name + ".py", "exec"), mod.__dict__)
    return mod


@pytest.mark.parametrize("module", sorted(FUNCS))
def test_public_signatures_equal_the_reference(module):
    if module != "utils":
        _load_reference_module("utils")
    ref = _load_reference_module(module)
    ours = importlib.import_module(f"gaussianimage_plus_b200.gsplat.{module}")
    for fn in FUNCS[module]:
        want = inspect.signature(getattr(ref, fn))
        got = inspect.signature(getattr(ours, fn))
        wp, gp = list(want.parameters.values()), list(got.parameters.values())
        if fn == "rasterize_gaussians_sum":
            # the clean 15-parameter form is accepted positionally; the compat shim takes *args
            assert any(p.kind == p.VAR_POSITIONAL for p in gp) or [p.name for p in gp] == [p.name for p in wp]
            continue
        # extensions must be keyword-only or underscore-prefixed trailing parameters with defaults
        names_w = [p.name for p in wp]
        names_g = [p.name for p in gp if not p.name.startswith("_")]
        assert names_g[:len(names_w)] == names_w, (fn, names_g, names_w)
        for extra in gp[len(wp):]:
            assert extra.default is not inspect.Parameter.empty, (fn, extra.name)
        for a, b in zip(wp, gp):
            assert (a.default is inspect.Parameter.empty) == (b.default is inspect.Parameter.empty), (fn, a.name)
            if a.default is not inspect.Parameter.empty:
                assert a.default == b.default, (fn, a.name, a.default, b.default)


def test_package_exports_what_the_reference_package_exports():
    import gaussianimage_plus_b200.gsplat as ours

    src = open(os.path.join(REF, "gsplat", "gsplat", "__init__.py")).read()
    for line in src.splitlines():
        line = line.split("#")[0].strip()
        if line.startswith("from .") and " import " in line and "version" not in line:
            for name in line.split(" import ")[1].replace("(", "").replace(")", "").split(","):
                name = name.strip()
                if name in ("project_gaussians", "rasterize_gaussians", "spherical_harmonics", "ProjectGaussians",
                            "RasterizeGaussians", "NDRasterizeGaussians", "SphericalHarmonics", "rasterize_gaussians_indices",
                            ""):
                    continue      # the 3-D pipeline: out of scope
                if name.startswith(("project_gaussians_2d", "rasterize_gaussians_sum", "rasterize_gaussians_plus")):
                    assert hasattr(ours, name), name


def test_reference_model_file_imports_on_the_dropin():
    """models/gaussianimage_covariance.py of the reference, unmodified, with `gsplat` = this package."""
    import gaussianimage_plus_b200 as pkg

    saved = {k: sys.modules.get(k) for k in ("gsplat", "utils", "quantize", "optimizer", "models", "models.utils")}
    stubs = []
    try:
        pkg.install_as_gsplat()
        for name in ("vector_quantize_pytorch", "constriction", "pytorch_msssim", "matplotlib", "matplotlib.pyplot",
                     "matplotlib.patches", "lpips", "cv2"):
            if name not in sys.modules:
                m = types.ModuleType(name)
                m.__path__ = []            # (a package, so that `import matplotlib.patches` resolves to the stub)
                m.VectorQuantize = m.ResidualVQ = m.Ellipse = object
                m.ms_ssim = m.ssim = lambda *a, **k: None
                sys.modules[name] = m
                stubs.append(name)
        sys.path.insert(0, REF)
        for k in ("utils", "quantize", "optimizer", "models", "models.utils"):
            sys.modules.pop(k, None)
        spec = importlib.util.spec_from_file_location("_ref_model", os.path.join(REF, "models", "gaussianimage_covariance.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert hasattr(mod, "GaussianImage_Covariance")
        # the operators the model file bound at import time are this package's
        import gaussianimage_plus_b200.gsplat as ours

        assert mod.project_gaussians_2d_covariance is ours.project_gaussians_2d_covariance
        assert mod.rasterize_gaussians_plus is ours.rasterize_gaussians_plus
        # and the forward() source still calls them with the arguments the signatures accept
        src = inspect.getsource(mod.GaussianImage_Covariance.forward)
        assert "project_gaussians_2d_covariance(" in src and "rasterize_gaussians_plus(" in src
    finally:
        if REF in sys.path:
            sys.path.remove(REF)
        for name in stubs:
            sys.modules.pop(name, None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
