"""ctypes/numpy front-end of the CPU ORACLE (oracle/dm2_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference fallback).
Never imported by dmesh2_renderer_b200/.

The oracle works at the same boundary as the reference's pybind functions: it takes the 21
``RenderFunction`` inputs (numpy, host memory) and returns colour/depth, the per-pixel state, the
binning taps (keys, sorted list, ranges) and fp64-accumulated gradients.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "dm2_oracle.cpp")
LIB = os.path.join(_HERE, "_build", "libdm2_oracle.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                           "-fno-fast-math", SRC, "-o", LIB])
    return LIB


class OrcRenderIn(C.Structure):
    _fields_ = [("B", C.c_int32), ("P", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("aa_temperature", C.c_float), ("K", C.c_int32), ("nthreads", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("background", "patch_min", "verts", "faces", "verts_color", "faces_opacity",
                                          "verts_ndc", "verts_image", "faces_intense", "aa_verts", "aa_edges",
                                          "aa_iszero", "aa_recip", "aa_normal", "aa_normal_c", "ray_o", "ray_d")]


class OrcLayerIn(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "P", "F", "T", "H", "W", "L", "nthreads")] + \
               [(n, C.c_void_p) for n in ("verts", "faces", "tets", "face_tets", "tet_faces", "face_existence",
                                          "ray_o", "ray_d")]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_hardware_threads.restype = C.c_int
        _lib.orc_higher_msb.restype = C.c_uint32
        _lib.orc_higher_msb.argtypes = [C.c_uint32]
        _lib.orc_bin.restype = C.c_longlong
        _lib.orc_aa_overlap.restype = C.c_int
        _lib.orc_aa_overlap.argtypes = [C.c_void_p] * 6 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    return _lib


def hardware_threads():
    return int(lib().orc_hardware_threads())


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _c(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def higher_msb(n):
    return int(lib().orc_higher_msb(int(n)))


def preprocess(B, P, F, H, W, patch_min, faces, verts_ndc, verts_image):
    L = lib()
    n = B * F
    out = dict(depths=np.full(n, np.nan, np.float32), min_depths=np.full(n, np.nan, np.float32),
               max_depths=np.full(n, np.nan, np.float32), tiles_touched=np.zeros(n, np.uint32),
               rects=np.zeros((n, 4), np.int32))
    pm = _c(patch_min, np.int32) if patch_min is not None else None
    L.orc_preprocess(C.c_int(B), C.c_int(P), C.c_int(F), C.c_int(H), C.c_int(W), _p(pm), _p(faces), _p(verts_ndc),
                     _p(verts_image), _p(out["depths"]), _p(out["min_depths"]), _p(out["max_depths"]),
                     _p(out["tiles_touched"]), _p(out["rects"]))
    out["face_offsets"] = np.cumsum(out["tiles_touched"].astype(np.uint64)).astype(np.uint32)
    return out


def bin_faces(B, F, H, W, key_depths, tiles_touched, rects):
    L = lib()
    args = [C.c_int(B), C.c_int(F), C.c_int(H), C.c_int(W), _p(key_depths), _p(tiles_touched), _p(rects)]
    R = int(L.orc_bin(*args, None, None, None, None, None))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out = dict(num_rendered=R, keys_unsorted=np.zeros(R, np.uint64), vals_unsorted=np.zeros(R, np.uint32),
               keys=np.zeros(R, np.uint64), vals=np.zeros(R, np.uint32), ranges=np.zeros((B * tiles, 2), np.uint32))
    L.orc_bin(*args, _p(out["keys_unsorted"]), _p(out["vals_unsorted"]), _p(out["keys"]), _p(out["vals"]),
              _p(out["ranges"]))
    return out


RENDER_KEYS = ("background", "patch_min", "verts", "faces", "verts_color", "faces_opacity", "verts_ndc",
               "verts_image", "faces_intense", "aa_face_verts", "aa_face_edges", "aa_face_edges_iszero",
               "aa_face_edges_recip", "aa_face_edges_normal", "aa_face_edges_normal_c", "image_ray_o", "image_ray_d")


def _render_struct(a, patch_width, patch_height, aa_temperature, K, nthreads):
    keep = dict(
        background=_c(a["background"], np.float32), patch_min=_c(a["patch_min"], np.int32),
        verts=_c(a["verts"], np.float32), faces=_c(a["faces"], np.int32),
        verts_color=_c(a["verts_color"], np.float32), faces_opacity=_c(a["faces_opacity"], np.float32),
        verts_ndc=_c(a["verts_ndc"], np.float32), verts_image=_c(a["verts_image"], np.float32),
        faces_intense=_c(a["faces_intense"], np.float32), aa_verts=_c(a["aa_face_verts"], np.float32),
        aa_edges=_c(a["aa_face_edges"], np.float32), aa_iszero=_c(a["aa_face_edges_iszero"], np.uint8),
        aa_recip=_c(a["aa_face_edges_recip"], np.float32), aa_normal=_c(a["aa_face_edges_normal"], np.float32),
        aa_normal_c=_c(a["aa_face_edges_normal_c"], np.float32), ray_o=_c(a["image_ray_o"], np.float32),
        ray_d=_c(a["image_ray_d"], np.float32))
    s = OrcRenderIn()
    s.B, s.P, s.F = keep["verts_ndc"].shape[0], keep["verts"].shape[0], keep["faces"].shape[0]
    s.H, s.W = int(patch_height), int(patch_width)
    s.aa_temperature, s.K, s.nthreads = float(aa_temperature), int(K), int(nthreads)
    for k, v in keep.items():
        setattr(s, k, v.ctypes.data)
    return s, keep


def render(args, patch_width, patch_height, aa_temperature=1.0, K=0, nthreads=None, backward=True,
           dL_dcolor=None, dL_ddepth=None, forward_only_timing=False):
    """Full reference pipeline on the CPU.  ``args``: dict with RENDER_KEYS (numpy arrays).

    K = reference's len_oarea_buffer: 0 recomputes AA terms in backward (reference behaviour for
    K == 0, backward.cu:264-272); K > 0 reproduces the reference's per-pixel AA cache.
    """
    L = lib()
    nthreads = nthreads or hardware_threads()
    s, keep = _render_struct(args, patch_width, patch_height, aa_temperature, K, nthreads)
    B, P, F, H, W = s.B, s.P, s.F, s.H, s.W
    out = {}
    pre = preprocess(B, P, F, H, W, keep["patch_min"], keep["faces"], keep["verts_ndc"], keep["verts_image"])
    out.update(pre)
    b = bin_faces(B, F, H, W, pre["depths"], pre["tiles_touched"], pre["rects"])
    out.update(b)
    npix = B * H * W
    out["color"] = np.zeros((B, H, W, 3), np.float32)
    out["depth"] = np.zeros((B, H, W), np.float32)
    out["final_T"] = np.zeros((B, H, W), np.float32)
    out["final_prev_T"] = np.zeros((B, H, W), np.float32)
    out["n_contrib"] = np.zeros((B, H, W), np.uint32)
    Kc = 0 if aa_temperature == 0.0 else int(K)
    cache = None
    if Kc > 0:
        cache = dict(oarea=np.zeros((npix, Kc), np.float32), tri_id=np.zeros((npix, Kc), np.int32),
                     cnt=np.zeros(npix, np.int32), doarea=np.zeros((npix, Kc, 6), np.float32))
    cp = (lambda k: _p(cache[k])) if cache else (lambda k: None)
    L.orc_render_forward(C.byref(s), _p(b["vals"]), _p(b["ranges"]), _p(out["color"]), _p(out["depth"]),
                         _p(out["final_T"]), _p(out["final_prev_T"]), _p(out["n_contrib"]), cp("oarea"),
                         cp("tri_id"), cp("cnt"), cp("doarea"))
    if backward:
        gc = _c(dL_dcolor, np.float32)
        gd = _c(dL_ddepth, np.float32)
        g = dict(dL_dverts=np.zeros((P, 3), np.float64), dL_dverts_color=np.zeros((P, 3), np.float64),
                 dL_dfaces_opacity=np.zeros(F, np.float64), dL_dverts_ndc=np.zeros((B, P, 3), np.float64),
                 dL_dfaces_intense=np.zeros((B, F), np.float64), dL_daa_face_verts=np.zeros((B, F, 3, 2), np.float64))
        L.orc_render_backward(C.byref(s), _p(b["vals"]), _p(b["ranges"]), _p(out["final_T"]), _p(out["final_prev_T"]),
                              _p(out["n_contrib"]), _p(gc), _p(gd), cp("oarea"), cp("tri_id"), cp("cnt"), cp("doarea"),
                              _p(g["dL_dverts"]), _p(g["dL_dverts_color"]), _p(g["dL_dfaces_opacity"]),
                              _p(g["dL_dverts_ndc"]), _p(g["dL_dfaces_intense"]), _p(g["dL_daa_face_verts"]))
        out.update(g)
    return out


def layers(verts, faces, tets, face_tets, tet_faces, face_existence, verts_ndc, verts_image, ray_o, ray_d, width,
           height, num_layers, nthreads=None):
    L = lib()
    nthreads = nthreads or hardware_threads()
    keep = dict(verts=_c(verts, np.float32), faces=_c(faces, np.int32), tets=_c(tets, np.int32),
                face_tets=_c(face_tets, np.int32), tet_faces=_c(tet_faces, np.int32),
                face_existence=_c(face_existence, np.int32), ray_o=_c(ray_o, np.float32), ray_d=_c(ray_d, np.float32))
    vn, vi = _c(verts_ndc, np.float32), _c(verts_image, np.float32)
    B, P, F, T = vn.shape[0], keep["verts"].shape[0], keep["faces"].shape[0], keep["tets"].shape[0]
    H, W, Ln = int(height), int(width), int(num_layers)
    pre = preprocess(B, P, F, H, W, None, keep["faces"], vn, vi)
    b = bin_faces(B, F, H, W, pre["min_depths"], pre["tiles_touched"], pre["rects"])
    s = OrcLayerIn()
    s.B, s.P, s.F, s.T, s.H, s.W, s.L, s.nthreads = B, P, F, T, H, W, Ln, nthreads
    for k, v in keep.items():
        setattr(s, k, v.ctypes.data)
    out = dict(pre)
    out.update(b)
    out["first_face"] = np.full((B, H, W), -1, np.int32)
    out["first_tet"] = np.full((B, H, W), -1, np.int32)
    out["render_layers"] = np.full((B, H, W, Ln), -1, np.int32)
    out["render_layers_cnt"] = np.zeros((B, H, W), np.int32)
    L.orc_layers(C.byref(s), _p(pre["min_depths"]), _p(pre["max_depths"]), _p(b["vals"]), _p(b["ranges"]),
                 _p(out["first_face"]), _p(out["first_tet"]), _p(out["render_layers"]), _p(out["render_layers_cnt"]))
    return out


def aa_overlap(verts, edges, iszero, recip, normal, normal_c, pxmin, pymin):
    """One (triangle, unit pixel) pair.  Returns (error_code, area, grad[3,2])."""
    v, e, z = _c(verts, np.float32), _c(edges, np.float32), _c(iszero, np.uint8)
    r, n, c = _c(recip, np.float32), _c(normal, np.float32), _c(normal_c, np.float32)
    area = np.zeros(1, np.float32)
    g = np.zeros(6, np.float32)
    err = lib().orc_aa_overlap(_p(v), _p(e), _p(z), _p(r), _p(n), _p(c), C.c_float(pxmin), C.c_float(pymin),
                               _p(area), _p(g))
    return int(err), float(area[0]), g.reshape(3, 2)
