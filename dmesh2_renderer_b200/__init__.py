"""dmesh2_renderer_b200 - B200-native (sm_100a) drop-in for SonSang/dmesh2_renderer.

    import dmesh2_renderer_b200 as dmesh2_renderer
    r = dmesh2_renderer.Renderer(mv, proj, W, H, device)              # same ctor / forward signature
    lr = dmesh2_renderer.LayeredRenderer(mv, proj, W, H, device)      # same ctor / generate signature
    color, alpha = dmesh2_renderer.composite_layers(layers, cnt, ...)  # the consumer of generate() (not in the reference)

Host code is Python/PyTorch; all device work is hand-written CUDA in lib/libdm2_b200.so behind the
C-ABI of include/dm2_renderer.h.  There is no CPU fallback: without the built library every call
raises RuntimeError.
"""
from ._lib import launch_count, load_library, set_exact_clip   # noqa: F401
from .function import RenderFunction                   # noqa: F401
from .layers_composite import composite_layers, composite_layers_torch   # noqa: F401
from .renderer import LayeredRenderer, Renderer        # noqa: F401
from .triangles import Triangles                       # noqa: F401

__all__ = ["Renderer", "LayeredRenderer", "RenderFunction", "Triangles", "composite_layers", "composite_layers_torch",
           "load_library", "launch_count"]
