"""sinnerf_b200 -- Blackwell (sm_100a) volumetric renderer behind SinNeRF's render_rays /
NeRF / Embedding interface.  See DESIGN.md and INTEGRATION.md."""
from .config import get_precision, set_precision  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch-dependent modules are imported lazily so `import sinnerf_b200.build` stays light
    if name in ("render_rays", "render_rays_multi", "sample_pdf", "eval_points"):
        from . import rendering
        return getattr(rendering, name)
    if name in ("NeRF", "Embedding"):
        from . import nerf
        return getattr(nerf, name)
    raise AttributeError(name)
