"""ctypes binding of libdm2_b200.so (include/dm2_renderer.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU or
PyTorch fallback: if the shared object is missing, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdm2_b200.so")
ABI_VERSION = 17
RAY_RECIPES = 36        # DM2_RAY_RECIPES

c_f32p = C.c_void_p   # all device pointers travel as integers


class DM2RenderInputs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("P", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("aa_temperature", C.c_float), ("len_oarea_buffer", C.c_int32),
        ("background", C.c_void_p), ("patch_min", C.c_void_p), ("verts", C.c_void_p), ("faces", C.c_void_p),
        ("verts_color", C.c_void_p), ("faces_opacity", C.c_void_p), ("verts_ndc", C.c_void_p),
        ("verts_image", C.c_void_p), ("faces_intense", C.c_void_p), ("aa_face_verts", C.c_void_p),
        ("aa_face_edges", C.c_void_p), ("aa_face_edges_iszero", C.c_void_p), ("aa_face_edges_recip", C.c_void_p),
        ("aa_face_edges_normal", C.c_void_p), ("aa_face_edges_normal_c", C.c_void_p),
        ("image_ray_o", C.c_void_p), ("image_ray_d", C.c_void_p),
        ("ray_view", C.c_void_p), ("ray_h", C.c_int32), ("ray_w", C.c_int32),
        ("depth_unit", C.c_int32), ("ray_recipe", C.c_int32), ("ray_params", C.c_void_p),
        ("ray_v", C.c_int32), ("reserved_", C.c_int32),
    ]


class DM2FusedInputs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("P", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("image_h", C.c_int32), ("image_w", C.c_int32), ("aa_temperature", C.c_float),
        ("background", C.c_void_p), ("patch_min", C.c_void_p), ("verts", C.c_void_p), ("faces", C.c_void_p),
        ("verts_color", C.c_void_p), ("faces_opacity", C.c_void_p), ("mv", C.c_void_p), ("proj", C.c_void_p),
        ("faces_intense", C.c_void_p), ("ray_o", C.c_void_p), ("ray_d", C.c_void_p), ("ray_view", C.c_void_p),
        ("depth_unit", C.c_int32), ("ray_recipe", C.c_int32), ("ray_params", C.c_void_p),
        ("ray_v", C.c_int32), ("reserved_", C.c_int32),
    ]


class DM2FusedGrads(C.Structure):
    _fields_ = [
        ("dL_dverts", C.c_void_p), ("dL_dverts_color", C.c_void_p), ("dL_dfaces_opacity", C.c_void_p),
        ("dL_dfaces_intense", C.c_void_p),
    ]


class DM2RenderGrads(C.Structure):
    _fields_ = [
        ("dL_dverts", C.c_void_p), ("dL_dverts_color", C.c_void_p), ("dL_dfaces_opacity", C.c_void_p),
        ("dL_dverts_ndc", C.c_void_p), ("dL_dfaces_intense", C.c_void_p), ("dL_daa_face_verts", C.c_void_p),
    ]


class DM2LayerInputs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("P", C.c_int32), ("F", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("num_layers", C.c_int32),
        ("verts", C.c_void_p), ("faces", C.c_void_p), ("tets", C.c_void_p), ("face_tets", C.c_void_p),
        ("tet_faces", C.c_void_p), ("face_existence", C.c_void_p), ("verts_ndc", C.c_void_p),
        ("verts_image", C.c_void_p), ("image_ray_o", C.c_void_p), ("image_ray_d", C.c_void_p),
        ("mv", C.c_void_p), ("proj", C.c_void_p), ("ray_view", C.c_void_p),
        ("ray_params", C.c_void_p), ("ray_recipe", C.c_int32), ("ray_v", C.c_int32),
    ]


class DM2FaceLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "depths", "min_depths", "max_depths", "tiles_touched", "rects",
        "records", "record_stride", "total")]


class DM2ImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "ranges", "tile_cursor", "final_T", "final_prev_T", "n_contrib", "first_face", "first_tet", "total")]


class DM2BinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("keys_unsorted", "vals_unsorted", "keys", "vals", "sort_temp", "imask", "meta", "total")]


# every symbol declared in include/dm2_renderer.h: (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "dm2_last_error": (C.c_char_p, []),
    "dm2_abi_version": (C.c_int, []),
    "dm2_face_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P(DM2FaceLayout)]),
    "dm2_image_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P(DM2ImageLayout)]),
    "dm2_binning_layout": (C.c_int, [C.c_int64, _P(DM2BinningLayout)]),
    "dm2_render_preprocess": (C.c_int, [_P(DM2RenderInputs), C.c_void_p, C.c_void_p, _P(C.c_int64), _P(C.c_int64),
                                        C.c_void_p]),
    "dm2_render_forward": (C.c_int, [_P(DM2RenderInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dm2_backward_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, _P(C.c_size_t)]),
    "dm2_render_backward": (C.c_int, [_P(DM2RenderInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, _P(DM2RenderGrads), C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p]),
    "dm2_fused_preprocess": (C.c_int, [_P(DM2FusedInputs), C.c_void_p, C.c_void_p, _P(C.c_int64), _P(C.c_int64),
                                       C.c_void_p]),
    "dm2_fused_forward": (C.c_int, [_P(DM2FusedInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dm2_fused_forward_pipelined": (C.c_int, [_P(DM2FusedInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              _P(C.c_int64), _P(C.c_int64), _P(C.c_int32), C.c_void_p]),
    "dm2_fused_backward": (C.c_int, [_P(DM2FusedInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, _P(DM2FusedGrads), C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p]),
    "dm2_generate_rays": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "dm2_peer_allreduce_sum": (C.c_int, [_P(C.c_void_p), C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "dm2_peer_allreduce_sum_fused": (C.c_int, [_P(C.c_void_p), C.c_void_p, _P(C.c_void_p), C.c_int32, C.c_int32,
                                               C.c_int64, C.c_uint32, C.c_void_p]),
    "dm2_layers_preprocess": (C.c_int, [_P(DM2LayerInputs), C.c_void_p, C.c_void_p, _P(C.c_int64), _P(C.c_int64),
                                        C.c_void_p]),
    "dm2_layers_geo_bytes": (C.c_int, [C.c_int32, C.c_int32, _P(C.c_size_t)]),
    "dm2_generate_render_layers": (C.c_int, [_P(DM2LayerInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dm2_layers_preprocess_with_geometry": (C.c_int, [_P(DM2LayerInputs), C.c_void_p, C.c_void_p, C.c_void_p,
                                                      _P(C.c_int64), _P(C.c_int64), C.c_void_p]),
    "dm2_generate_render_layers_prepared": (C.c_int, [_P(DM2LayerInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dm2_regenerate_render_layers": (C.c_int, [_P(DM2LayerInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p]),
    "dm2_layers_composite_forward": (C.c_int, [C.c_int32] * 5 + [C.c_void_p] * 9),
    "dm2_layers_composite_backward": (C.c_int, [C.c_int32] * 5 + [C.c_void_p] * 13),
    "dm2_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "dm2_sort_temp_bytes": (C.c_int, [C.c_int64, _P(C.c_size_t)]),
    "dm2_inclusive_scan_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dm2_scan_partials_count": (C.c_int, [C.c_int64, _P(C.c_int64)]),
    "dm2_higher_msb": (C.c_uint32, [C.c_uint32]),
    "dm2_launch_count": (C.c_int64, [C.c_int32]),
    "dm2_profile_enable": (None, [C.c_int32]),
    "dm2_set_forward_rounds": (None, [C.c_int64, C.c_int64]),
    "dm2_set_exact_clip": (None, [C.c_int32]),
    "dm2_get_exact_clip": (C.c_int32, []),
    "dm2_profile_reset": (None, []),
    "dm2_profile_report": (C.c_int64, [C.c_char_p, C.c_int64]),
}

_lib = None


def load_library(path=None):
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("DM2_LIB") or LIB_PATH      # DM2_LIB: A/B builds of tools/build_variant.py
    if not os.path.exists(p):
        raise RuntimeError(
            f"dmesh2_renderer_b200: CUDA library {p} not found; run `python __graft_entry__.py` "
            "(nvcc, sm_100a) first.  There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError = symbol missing = broken build
        fn.restype = res
        fn.argtypes = args
    if lib.dm2_abi_version() != ABI_VERSION:
        raise RuntimeError(f"dmesh2_renderer_b200: ABI mismatch ({lib.dm2_abi_version()} != {ABI_VERSION})")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load_library().dm2_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"libdm2_b200 call failed with status {status}")


def set_exact_clip(on=True):
    """True: every partially covered (pixel, face) pair takes the reference's polygon construction (forward images
    bit-identical to the reference's); False (default): the polygon-free overlap area of csrc/dm2_aa_fast.cuh where it
    yields the same value to ~1e-7 (pairs far from every rounding-dependent decision of the reference, thresholds
    measured against the oracle: tests/test_aa_fast.py), the construction elsewhere.  Process-wide.  Returns the previous setting."""
    lib = load_library()
    prev = bool(lib.dm2_get_exact_clip())
    lib.dm2_set_exact_clip(1 if on else 0)
    return prev


def profile_enable(on=True):
    load_library().dm2_profile_enable(1 if on else 0)


def profile_reset():
    load_library().dm2_profile_reset()


def profile_report():
    """{stage: (total_ms, launches)} measured with CUDA events on the launching stream."""
    lib = load_library()
    n = lib.dm2_profile_report(None, 0)
    buf = C.create_string_buffer(int(n) + 16)
    lib.dm2_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.rsplit(" ", 2)
        out[name] = (float(ms), int(cnt))
    return out


def launch_count(reset=False):
    return int(load_library().dm2_launch_count(1 if reset else 0))
