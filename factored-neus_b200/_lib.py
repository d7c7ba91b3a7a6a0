"""ctypes binding of libfneus_b200.so (C ABI: include/fneus.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing or a call
fails, the binding raises.  ``build_library()`` compiles it in-tree with nvcc
for sm_100a (``__graft_entry__.build`` calls it).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfneus_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class SdfCfg(Structure):
    _fields_ = [("d_in", c_int), ("d_hidden", c_int), ("n_layers", c_int), ("d_out", c_int),
                ("multires", c_int), ("skip_layer", c_int), ("scale", c_float), ("beta", c_float), ("precision", c_int),
                ("feat_image", c_int)]


class ColorCfg(Structure):
    _fields_ = [("d_feature", c_int), ("d_hidden", c_int), ("n_layers", c_int), ("d_out", c_int),
                ("multires_view", c_int), ("precision", c_int), ("feat_image", c_int)]


class RefCfg(Structure):
    _fields_ = [("d_feature", c_int), ("d_hidden", c_int), ("precision", c_int)]


class MlpCfg(Structure):
    _fields_ = [("n_inputs", c_int), ("in_dim", c_int * 2), ("in_multires", c_int * 2), ("d_hidden", c_int),
                ("n_layers", c_int), ("d_out", c_int), ("last_act", c_int), ("precision", c_int)]


class NerfCfg(Structure):
    _fields_ = [("D", c_int), ("W", c_int), ("d_in", c_int), ("d_in_view", c_int), ("multires", c_int),
                ("multires_view", c_int), ("skip", c_int), ("precision", c_int)]


def source_digest() -> str:
    """SHA-256 over every source the library is compiled from (csrc/*, include/fneus.h) and the compiler flags."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    for p in srcs + [os.path.join(os.path.dirname(_HERE), "include", "fneus.h")]:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> factored-neus_b200/libfneus_b200.so

    The library is rebuilt whenever the digest of its sources differs from the one recorded next to it at the last
    build (``libfneus_b200.so.digest``; file times are not trusted: a fresh clone or snapshot resets them)."""
    digest = source_digest()
    stamp = LIB_PATH + ".digest"
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "libfneus.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    with open(stamp, "w") as f:
        f.write(digest + "\n")
    return LIB_PATH


_P = c_void_p
_LL = c_longlong
_SIGNATURES = {
    "fneus_status_string": (ctypes.c_char_p, [c_int]),
    "fneus_abi_version": (c_int, []),
    "fneus_num_sms": (c_int, []),
    "fneus_surface_blend_fwd": (c_int, [_P, _P, _P, _P, _P, _LL, _P, _P, _P, _P]),
    "fneus_surface_blend_bwd": (c_int, [_P, _P, _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fneus_loss_norms": (c_int, [_P, _P, _P, _LL, c_int, _P, _P]),
    "fneus_stage1_loss": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _LL, c_int, c_float, c_float, c_float, _P, _P, _P,
                                  _P, _P, _P]),
    "fneus_lvis_trace_workspace_floats": (_LL, [POINTER(SdfCfg), POINTER(ColorCfg), _LL, c_int, c_int]),
    "fneus_lvis_trace": (c_int, [POINTER(SdfCfg), _P, POINTER(ColorCfg), _P, _P, _P, _LL, c_int, c_int, c_int, _P, _P, _P,
                                 _P, _P, _P, _P, _LL, _LL, _P]),
    "fneus_mc_classify": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    "fneus_mc_emit": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fneus_stage2_loss": (c_int, [_P, _P, _P, _P, _P, _P, _LL, c_int, _P, _P, _P, _P]),
    "fneus_gen_rays": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _LL, _P, _P, _P, _P]),
    "fneus_near_far": (c_int, [_P, _P, _LL, _P, _P, _P]),
    "fneus_coarse_z": (c_int, [_P, _P, _P, _P, _LL, c_int, c_float, _P, _P]),
    "fneus_hit_rows": (c_int, [_P, _LL, c_int, _P, _P]),
    "fneus_adam_step": (c_int, [_P, _P, _P, _P, _LL, _P, c_float, c_float, c_float, c_float, c_float, c_float, c_float,
                                c_float, c_int, _P]),
    "fneus_debug_flags": (c_int, [c_int]),
    "fneus_debug_hang_record": (c_int, [_P]),
    "fneus_debug_timeline": (c_int, [_P, c_int]),
    "fneus_debug_gemm": (c_int, [c_int, c_int, _P, c_int, _P, c_int, _P, _LL, c_int, c_int, _P, c_int, _P]),
    "fneus_prof_classes": (c_int, []),
    "fneus_prof_enable": (c_int, [c_int]),
    "fneus_prof_collect": (c_int, [_P, _P, _P, _P]),
    "fneus_upsample_iter": (c_int, [_P, _P, _P, _P, _LL, c_int, _P, _P, c_int, c_int, c_float, _P, _P, _P, _P, _P, _P]),
    "fneus_split_batch": (c_int, [_P, _LL, _P, _P, _P, _P, _P]),
    "fneus_inv_s": (c_int, [_P, _P, _P, _P]),
    "fneus_composite_post": (c_int, [_P, _P, _LL, _P, _P, _P]),
    "fneus_gather_rows3": (c_int, [_P, _P, _P, _P, _LL, _P, _P, _P, _P]),
    "fneus_scatter_rows3": (c_int, [_P, _P, _LL, _P, _P]),
    "fneus_image_bytes": (_LL, [_LL, c_int]),
    "fneus_image_gather_rows": (c_int, [_P, c_int, c_int, _P, _LL, _P, _P]),
    "fneus_image_scatter_add_rows": (c_int, [_P, c_int, _P, _LL, _P, _P]),
    "fneus_sdf_feat_image_ok": (c_int, [POINTER(SdfCfg)]),
    "fneus_color_feat_image_ok": (c_int, [POINTER(ColorCfg)]),
    "fneus_pack_max_segments": (c_int, []),
    "fneus_pack_fwd": (c_int, [c_int, _P, _P, _P, _P, _P, _P, _P]),
    "fneus_pack_bwd": (c_int, [c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "fneus_sdf_pack_floats": (_LL, [POINTER(SdfCfg)]),
    "fneus_sdf_saved_floats": (_LL, [POINTER(SdfCfg), _LL]),
    "fneus_sdf_scratch_floats": (_LL, [POINTER(SdfCfg), _LL]),
    "fneus_sdf_fwd": (c_int, [POINTER(SdfCfg), _P, _P, _LL, _P, _P, _P, _LL, _P]),
    "fneus_sdf_grid": (c_int, [POINTER(SdfCfg), _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _LL, _P]),
    "fneus_sdf_fwd_grad": (c_int, [POINTER(SdfCfg), _P, _P, _LL, _P, _P, _P, _P, _P, _P]),
    "fneus_sdf_bwd": (c_int, [POINTER(SdfCfg), _P, _P, _LL, _P, _P, _P, _P, _P, _P, _P]),
    "fneus_color_pack_floats": (_LL, [POINTER(ColorCfg)]),
    "fneus_color_saved_floats": (_LL, [POINTER(ColorCfg), _LL]),
    "fneus_color_scratch_floats": (_LL, [POINTER(ColorCfg), _LL]),
    "fneus_color_fwd": (c_int, [POINTER(ColorCfg), _P, _P, _P, _P, _P, _LL, _P, _P, _P, _P]),
    "fneus_color_bwd": (c_int, [POINTER(ColorCfg), _P, _P, _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fneus_ref_pack_floats": (_LL, [POINTER(RefCfg)]),
    "fneus_ref_saved_floats": (_LL, [POINTER(RefCfg), _LL]),
    "fneus_ref_scratch_floats": (_LL, [POINTER(RefCfg), _LL]),
    "fneus_ref_fwd": (c_int, [POINTER(RefCfg), _P, _P, _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P]),
    "fneus_ref_bwd": (c_int, [POINTER(RefCfg), _P, _P, _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fneus_mlp_pack_floats": (_LL, [POINTER(MlpCfg)]),
    "fneus_mlp_saved_floats": (_LL, [POINTER(MlpCfg), _LL]),
    "fneus_mlp_scratch_floats": (_LL, [POINTER(MlpCfg), _LL]),
    "fneus_mlp_fwd": (c_int, [POINTER(MlpCfg), _P, _P, _P, _LL, _P, _P, _P, _P]),
    "fneus_mlp_bwd": (c_int, [POINTER(MlpCfg), _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P]),
    "fneus_nerf_pack_floats": (_LL, [POINTER(NerfCfg)]),
    "fneus_nerf_saved_floats": (_LL, [POINTER(NerfCfg), _LL]),
    "fneus_nerf_scratch_floats": (_LL, [POINTER(NerfCfg), _LL]),
    "fneus_nerf_fwd": (c_int, [POINTER(NerfCfg), _P, _P, _P, _LL, _P, _P, _P, _P]),
    "fneus_nerf_bwd": (c_int, [POINTER(NerfCfg), _P, _P, _P, _LL, _P, _P, _P, _P, _P, _P]),
    "fneus_outside_geometry": (c_int, [_P, _P, _P, _LL, c_int, c_float, _P, _P, _P, _P]),
    "fneus_outside_alpha_fwd": (c_int, [_P, _P, _P, _LL, _P, _P, _P]),
    "fneus_outside_alpha_bwd": (c_int, [_P, _P, _P, _P, _P, _LL, _P, _P, _P]),
    "fneus_ray_points": (c_int, [_P, _P, _P, _LL, c_int, _P, _P]),
    "fneus_upsample_step": (c_int, [_P, _P, _P, _P, _LL, c_int, c_int, c_float, _P, _P, _P, _P, _P]),
    "fneus_upsample_step_dev": (c_int, [_P, _P, _P, _P, _LL, c_int, c_int, _P, _P, _P, _P]),
    "fneus_first_hit_secant": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _LL, c_int, _P, _P, _P, _P, _P, _P]),
    "fneus_inverse_cdf": (c_int, [_P, _P, _P, _LL, c_int, c_int, _P, _P, _P]),
    "fneus_merge_sorted": (c_int, [_P, _P, _P, _P, _LL, c_int, c_int, _P, _P, _P]),
    "fneus_core_geometry": (c_int, [_P, _P, _P, _LL, c_int, c_float, _P, _P, _P, _P, _P]),
    "fneus_composite_fwd": (c_int, [_P] * 9 + [_LL, c_int, c_int, _P, c_float] + [_P] * 11),
    "fneus_composite_bwd": (c_int, [_P] * 9 + [_LL, c_int, c_int, _P, c_float] + [_P] * 15),
}

_lib = None


def declared_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "factored-neus_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/PyTorch fallback for the CUDA path)" % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str = "fneus call"):
    if rc != 0:
        msg = lib().fneus_status_string(int(rc)).decode()
        raise RuntimeError("%s failed: %s (status %d)" % (what, msg, rc))


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "fneus: tensor must be contiguous"
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
