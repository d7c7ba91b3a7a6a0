"""torch.autograd.Function wrappers around the C ABI (include/fneus.h).

Each Function allocates outputs/workspaces with torch (device memory + stream
plumbing only) and hands raw pointers to libfneus_b200.so; the arithmetic is all
in the CUDA library.  Gradients are returned for the flat *effective* weight packs;
weight-norm stays in front as differentiable torch ops.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("factored-neus_b200: %s must be a CUDA tensor (no CPU fallback)" % name)


_IMG_SLACK_FLOATS = 1 << 20      # room for the BF16 weight images of one SDF pass (tensor-core mode)


def _empty(n, like):
    return torch.empty(int(n), dtype=torch.float32, device=like.device)


# ---------------------------------------------------------------------------------------------
# precision switch + raw GEMM test hook
# ---------------------------------------------------------------------------------------------
_PRECISION = {"default": 0}
_LIVE_MODULES = __import__("weakref").WeakSet()      # network modules whose cfg carries a precision field


def _prec_code(mode):
    return {"fp32": 0, "bf16": 1, "tc": 1, 0: 0, 1: 1}[mode]


def register_module(m):
    """Networks register themselves so that ``set_precision`` can retarget them; their cfg starts at the default."""
    m.cfg.precision = _PRECISION["default"]
    _LIVE_MODULES.add(m)


def set_precision(mode, modules=None):
    """'fp32' (CUDA-core exactness anchor, default) or 'bf16' / 'tc' (tcgen05 tensor cores: FP16 forward / BF16 backward
    operands, FP32 accumulate).  The precision is a field of each network's configuration struct -- the C library keeps no
    global mode.  With ``modules`` only those networks are retargeted (two renderers of one process may differ);
    without, the default for new networks AND every live network are set (convenience for scripts and tests)."""
    code = _prec_code(mode)
    if modules is None:
        _PRECISION["default"] = code
        modules = list(_LIVE_MODULES)
    for m in modules:
        m.cfg.precision = code


def get_precision():
    return "bf16" if _PRECISION["default"] == 1 else "fp32"


def debug_gemm(kind, A, W, bias=None, out=None):
    """kind 0: A[M,K] @ W[N,K].T + bias ; 1: A[M,K] @ W[K,N] ; 2: out[N,K] += W(=Y)[M,N].T @ A[M,K], bias += colsum(Y).
    A (and Y) need a leading dimension that is a multiple of 4."""
    assert A.stride(1) == 1 and W.stride(1) == 1 and A.dtype == torch.float32 and W.dtype == torch.float32
    M, K = A.shape
    if kind == 0:
        N = W.shape[0]
        C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    elif kind == 1:
        N = W.shape[1]
        C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    else:
        N = W.shape[1]
        C = out if out is not None else torch.zeros(N, K, dtype=torch.float32, device=A.device)
    L.check(L.lib().fneus_debug_gemm(_PRECISION["default"], kind, A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), L.ptr(bias), M, N, K,
                                     L.ptr(C), C.stride(0), L.stream_ptr()), "fneus_debug_gemm")
    return C


# ---------------------------------------------------------------------------------------------
# flat weight packs (weight-norm fused), one launch forward and one backward
# ---------------------------------------------------------------------------------------------
class PackWeights(torch.autograd.Function):
    """(layout, *params) -> flat FP32 pack.  ``layout`` = list of (kind, i_g, i_v, rows, cols) over ``params``:
    kind 'wn' -> W = g*v/|v|_row (torch weight_norm, fields.py:67-68), 'copy' -> the tensor itself.

    Backward writes dg/dv/db in one launch and returns the gradients to autograd like any other Function (hooks,
    ``torch.autograd.grad`` and DDP reducers see them).

    Direct accumulation is an explicit OPT-IN: parameters registered with ``parallel.GradBucket(direct=True)`` carry
    ``_fneus_direct_grad`` and own a ``.grad`` view into the flat bucket; only when EVERY trainable parameter of the
    pack is registered that way does the kernel accumulate straight into those buffers (autograd then receives
    ``None``: one AccumulateGrad kernel per parameter saved).  In that mode tensor hooks, post-accumulate hooks,
    ``torch.autograd.grad`` and DDP are not supported for these parameters -- the bucket's all-reduce replaces them."""

    @staticmethod
    def forward(ctx, layout, *params):
        import ctypes
        dev = params[0].device
        _need_cuda(params[0], "parameters")
        n = len(layout)
        total = sum(r * c for (_, _, _, r, c) in layout)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        ps = [p.detach() for p in params]
        for p in ps:
            assert p.is_contiguous() and p.dtype == torch.float32
        PT = ctypes.c_void_p * n
        g = PT(*[(ps[ig].data_ptr() if kind == "wn" else None) for (kind, ig, iv, r, c) in layout])
        v = PT(*[ps[iv].data_ptr() for (kind, ig, iv, r, c) in layout])
        offs, o = [], 0
        for (_, _, _, r, c) in layout:
            offs.append(o)
            o += r * c
        off = (ctypes.c_longlong * n)(*offs)
        rows = (ctypes.c_int * n)(*[r for (_, _, _, r, c) in layout])
        cols = (ctypes.c_int * n)(*[c for (_, _, _, r, c) in layout])
        L.check(L.lib().fneus_pack_fwd(n, g, v, off, rows, cols, L.ptr(flat), L.stream_ptr()), "fneus_pack_fwd")
        ctx.layout, ctx.params, ctx.meta = layout, params, (g, v, off, rows, cols)
        return flat

    @staticmethod
    def backward(ctx, dflat):
        import ctypes
        layout, params = ctx.layout, ctx.params
        g, v, off, rows, cols = ctx.meta
        n = len(layout)
        dflat = _f32c(dflat)
        direct = all((getattr(p, "_fneus_direct_grad", False) and p.grad is not None and p.grad.is_contiguous())
                     or not p.requires_grad for p in params)
        grads = [None] * len(params)
        if not direct:
            grads = [torch.zeros_like(p) if p.requires_grad else None for p in params]
        tgt = (lambda i: (params[i].grad if direct else grads[i]) if params[i].requires_grad else None)
        PT = ctypes.c_void_p * n
        dg = PT(*[(tgt(ig).data_ptr() if (kind == "wn" and tgt(ig) is not None) else None)
                  for (kind, ig, iv, r, c) in layout])
        dv = PT(*[(tgt(iv).data_ptr() if tgt(iv) is not None else None) for (kind, ig, iv, r, c) in layout])
        L.check(L.lib().fneus_pack_bwd(n, g, v, dg, dv, off, rows, cols, L.ptr(dflat), 1 if direct else 0,
                                       L.stream_ptr()), "fneus_pack_bwd")
        return (None,) + tuple(None if direct else gr for gr in grads)


def pack_weights(layers):
    """layers: list of modules with (weight_g, weight_v, bias) or (weight, bias) -> flat pack [W0, b0, W1, b1, ...]."""
    params, layout = [], []
    for lin in layers:
        if hasattr(lin, "weight_g"):
            r, c = lin.weight_v.shape
            layout.append(("wn", len(params), len(params) + 1, r, c))
            params += [lin.weight_g, lin.weight_v]
        else:
            r, c = lin.weight.shape
            layout.append(("copy", -1, len(params), r, c))
            params += [lin.weight]
        layout.append(("copy", -1, len(params), 1, lin.bias.numel()))
        params += [lin.bias]
    return PackWeights.apply(layout, *params)


# ---------------------------------------------------------------------------------------------
# SDF network
# ---------------------------------------------------------------------------------------------
def sdf_forward_nograd(cfg, wflat, x, want_feat, max_chunk=1 << 18):
    """SDFNetwork.forward / .sdf without a graph (fields.py:74-95).  Returns (sdf [N,1], feat|None)."""
    _need_cuda(x, "x")
    x = _f32c(x)
    w = _f32c(wflat)
    N = x.shape[0]
    sdf = torch.empty(N, 1, dtype=torch.float32, device=x.device)
    feat = torch.empty(N, cfg.d_out - 1, dtype=torch.float32, device=x.device) if want_feat else None
    chunk = max(1, min(N, max_chunk))
    per_point = 2 * max(4, -(-cfg.d_hidden // 4) * 4, -(-(cfg.d_in * (1 + 2 * cfg.multires)) // 4) * 4,
                        -(-cfg.d_out // 4) * 4)
    scratch = _empty(per_point * chunk + _IMG_SLACK_FLOATS, x)
    L.check(L.lib().fneus_sdf_fwd(cfg, L.ptr(w), L.ptr(x), N, L.ptr(sdf), L.ptr(feat), L.ptr(scratch),
                                  scratch.numel(), L.stream_ptr()), "fneus_sdf_fwd")
    return sdf, feat


def sdf_grid(cfg, wflat, ax, ay, az, ix0=0, ix1=None, max_chunk=1 << 18):
    """-sdf on the ij-meshgrid of three axis tables (renderer.py:14-29), x-slab [ix0, ix1)."""
    _need_cuda(ax, "ax")
    w = _f32c(wflat)
    nx, ny, nz = ax.numel(), ay.numel(), az.numel()
    ix1 = nx if ix1 is None else ix1
    u = torch.empty(ix1 - ix0, ny, nz, dtype=torch.float32, device=ax.device)
    n = (ix1 - ix0) * ny * nz
    chunk = max(1, min(n, max_chunk))
    per_point = 2 * max(4, -(-cfg.d_hidden // 4) * 4, -(-(cfg.d_in * (1 + 2 * cfg.multires)) // 4) * 4,
                        -(-cfg.d_out // 4) * 4) + 3
    scratch = _empty(per_point * chunk + _IMG_SLACK_FLOATS, ax)
    L.check(L.lib().fneus_sdf_grid(cfg, L.ptr(w), L.ptr(_f32c(ax)), L.ptr(_f32c(ay)), L.ptr(_f32c(az)), nx, ny, nz,
                                   ix0, ix1, L.ptr(u), L.ptr(scratch), scratch.numel(), L.stream_ptr()),
            "fneus_sdf_grid")
    return u


def _image_cfg(cfg):
    """Copy of a network configuration with the feature hand-over switched to operand images (fneus.h: feat_image)."""
    c = type(cfg).from_buffer_copy(cfg)
    c.feat_image = 1
    return c


def feature_image_floats(n_rows, n_cols=256):
    """Size, in float32 words, of the opaque tensor that carries an [n_rows, n_cols] 16-bit operand image."""
    return int(L.lib().fneus_image_bytes(int(n_rows), int(n_cols))) // 4


def image_gather_rows(image, rows, n_cols=256, is_fp16=True):
    """FP32 rows [len(rows), n_cols] out of an operand image (no autograd; see GatherRows)."""
    out = torch.empty(rows.shape[0], n_cols, dtype=torch.float32, device=image.device)
    L.check(L.lib().fneus_image_gather_rows(L.ptr(image), 1 if is_fp16 else 0, n_cols, L.ptr(rows), rows.shape[0],
                                            L.ptr(out), L.stream_ptr()), "fneus_image_gather_rows")
    return out


class SdfValueGrad(torch.autograd.Function):
    """(wflat, x) -> (sdf [N,1], feat [N,d_out-1], normal [N,3]); fields.py:74-111 with create_graph=True
    semantics: all three outputs are differentiable w.r.t. the weights (double backward for the normal)."""

    @staticmethod
    def forward(ctx, wflat, x, cfg, want_normal, feat_image=False):
        _need_cuda(x, "x")
        xc, w = _f32c(x), _f32c(wflat)
        N = xc.shape[0]
        lib = L.lib()
        sdf = torch.empty(N, 1, dtype=torch.float32, device=x.device)
        if feat_image:
            # the features leave as an FP16 operand image inside an opaque float32 tensor; its gradient comes back as a
            # BF16 image in a tensor of the same shape (ColorMLP / FanOut / GatherRows know the convention)
            cfg = _image_cfg(cfg)
            feat = torch.empty(feature_image_floats(N, cfg.d_out - 1), dtype=torch.float32, device=x.device)
        else:
            feat = torch.empty(N, cfg.d_out - 1, dtype=torch.float32, device=x.device)
        normal = torch.empty(N, 3, dtype=torch.float32, device=x.device) if want_normal else None
        saved = _empty(lib.fneus_sdf_saved_floats(cfg, N), x)
        scratch = _empty(lib.fneus_sdf_scratch_floats(cfg, N), x)
        L.check(lib.fneus_sdf_fwd_grad(cfg, L.ptr(w), L.ptr(xc), N, L.ptr(sdf), L.ptr(feat), L.ptr(normal),
                                       L.ptr(saved), L.ptr(scratch), L.stream_ptr()), "fneus_sdf_fwd_grad")
        ctx.cfg, ctx.want_normal, ctx.consumed = cfg, want_normal, False
        ctx.save_for_backward(w, xc, saved)
        if not want_normal:
            normal = torch.zeros(0, 3, dtype=torch.float32, device=x.device)
            ctx.mark_non_differentiable(normal)
        return sdf, feat, normal

    # called (if set) when the SDF backward is about to start: everything autograd created AFTER the SDF forward -- the
    # colour / RefColor / variance gradients -- is final by then (parallel.GradBucket.all_reduce_segment uses it to send
    # that slice while the SDF backward runs)
    pre_backward_hook = None

    @staticmethod
    def backward(ctx, d_sdf, d_feat, d_normal):
        if SdfValueGrad.pre_backward_hook is not None and ctx.want_normal:
            SdfValueGrad.pre_backward_hook()
        if ctx.consumed:
            raise RuntimeError("fneus SdfValueGrad: backward twice is not supported (saved activations are "
                               "consumed in place)")
        ctx.consumed = True
        w, xc, saved = ctx.saved_tensors
        cfg = ctx.cfg
        N = xc.shape[0]
        lib = L.lib()
        dw = torch.zeros_like(w)
        scratch = _empty(lib.fneus_sdf_scratch_floats(cfg, N), xc)
        d_sdf, d_feat = _f32c(d_sdf), _f32c(d_feat)
        d_normal = _f32c(d_normal) if ctx.want_normal else None
        L.check(lib.fneus_sdf_bwd(cfg, L.ptr(w), L.ptr(xc), N, L.ptr(d_sdf), L.ptr(d_feat), L.ptr(d_normal),
                                  L.ptr(saved), L.ptr(scratch), L.ptr(dw), L.stream_ptr()), "fneus_sdf_bwd")
        return dw, None, None, None, None


# ---------------------------------------------------------------------------------------------
# colour network
# ---------------------------------------------------------------------------------------------
class ColorMLP(torch.autograd.Function):
    """RenderingNetwork.forward (fields.py:150-175): grads for weights, normals and features."""

    @staticmethod
    def forward(ctx, wflat, points, normals, view_dirs, feats, cfg, feat_image=False):
        _need_cuda(points, "points")
        w, p, n, v, f = _f32c(wflat), _f32c(points), _f32c(normals), _f32c(view_dirs), _f32c(feats)
        if feat_image:                                  # `feats` is SdfValueGrad's operand image
            cfg = _image_cfg(cfg)
        N = p.shape[0]
        lib = L.lib()
        rgb = torch.empty(N, cfg.d_out, dtype=torch.float32, device=p.device)
        need_graph = any(ctx.needs_input_grad)
        saved = _empty(lib.fneus_color_saved_floats(cfg, N), p) if need_graph else None
        scratch = _empty(lib.fneus_color_scratch_floats(cfg, N), p)
        L.check(lib.fneus_color_fwd(cfg, L.ptr(w), L.ptr(p), L.ptr(n), L.ptr(v), L.ptr(f), N, L.ptr(rgb),
                                    L.ptr(saved), L.ptr(scratch), L.stream_ptr()), "fneus_color_fwd")
        ctx.cfg = cfg
        if need_graph:
            ctx.save_for_backward(w, p, n, v, f, rgb, saved)
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        w, p, n, v, f, rgb, saved = ctx.saved_tensors
        cfg = ctx.cfg
        N = p.shape[0]
        lib = L.lib()
        dw = torch.zeros_like(w)
        d_n = torch.empty_like(n)
        d_f = torch.empty_like(f)
        scratch = _empty(lib.fneus_color_scratch_floats(cfg, N), p)
        L.check(lib.fneus_color_bwd(cfg, L.ptr(w), L.ptr(p), L.ptr(n), L.ptr(v), L.ptr(f), N, L.ptr(rgb),
                                    L.ptr(_f32c(d_rgb)), L.ptr(d_n), L.ptr(d_f), L.ptr(saved), L.ptr(scratch),
                                    L.ptr(dw), L.stream_ptr()), "fneus_color_bwd")
        return dw, None, d_n, None, d_f, None, None


class RefColorMLP(torch.autograd.Function):
    """RefColor.forward (fields.py:303-335) -> (rgb, specular_rgb, diffuse_rgb)."""

    @staticmethod
    def forward(ctx, wflat, points, feats, dirs, normals, cfg):
        _need_cuda(points, "points")
        w, p, f, d, n = _f32c(wflat), _f32c(points), _f32c(feats), _f32c(dirs), _f32c(normals)
        N = p.shape[0]
        lib = L.lib()
        outs = [torch.empty(N, 3, dtype=torch.float32, device=p.device) for _ in range(3)]
        saved = _empty(lib.fneus_ref_saved_floats(cfg, N), p)
        scratch = _empty(lib.fneus_ref_scratch_floats(cfg, N), p)
        L.check(lib.fneus_ref_fwd(cfg, L.ptr(w), L.ptr(p), L.ptr(f), L.ptr(d), L.ptr(n), N, L.ptr(outs[0]),
                                  L.ptr(outs[1]), L.ptr(outs[2]), L.ptr(saved), L.ptr(scratch), L.stream_ptr()),
                "fneus_ref_fwd")
        ctx.cfg = cfg
        ctx.save_for_backward(w, p, f, d, n, saved)
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_rgb, d_spec, d_diff):
        w, p, f, d, n, saved = ctx.saved_tensors
        cfg = ctx.cfg
        N = p.shape[0]
        lib = L.lib()
        dw = torch.zeros_like(w)
        d_f = torch.empty_like(f)
        d_n = torch.empty_like(n)
        scratch = _empty(lib.fneus_ref_scratch_floats(cfg, N), p)
        L.check(lib.fneus_ref_bwd(cfg, L.ptr(w), L.ptr(p), L.ptr(f), L.ptr(d), L.ptr(n), N, L.ptr(_f32c(d_rgb)),
                                  L.ptr(_f32c(d_spec)), L.ptr(_f32c(d_diff)), L.ptr(d_f), L.ptr(d_n),
                                  L.ptr(saved), L.ptr(scratch), L.ptr(dw), L.stream_ptr()), "fneus_ref_bwd")
        return dw, None, d_f, None, d_n, None


class PlainMLP(torch.autograd.Function):
    """Positional-encoded ReLU MLP without input gradients (Lvis, IndirectLight trunk): (wflat, in0, in1|None)."""

    @staticmethod
    def forward(ctx, wflat, in0, in1, cfg):
        _need_cuda(in0, "input")
        w, a, b = _f32c(wflat), _f32c(in0), _f32c(in1)
        N = a.shape[0]
        lib = L.lib()
        out = torch.empty(N, cfg.d_out, dtype=torch.float32, device=a.device)
        saved = _empty(lib.fneus_mlp_saved_floats(cfg, N), a)
        scratch = _empty(lib.fneus_mlp_scratch_floats(cfg, N), a)
        L.check(lib.fneus_mlp_fwd(cfg, L.ptr(w), L.ptr(a), L.ptr(b), N, L.ptr(out), L.ptr(saved), L.ptr(scratch),
                                  L.stream_ptr()), "fneus_mlp_fwd")
        ctx.cfg = cfg
        ctx.save_for_backward(w, a, b if b is not None else a, out, saved)
        ctx.has_b = b is not None
        return out

    @staticmethod
    def backward(ctx, d_out):
        w, a, b, out, saved = ctx.saved_tensors
        cfg = ctx.cfg
        N = a.shape[0]
        lib = L.lib()
        dw = torch.zeros_like(w)
        scratch = _empty(lib.fneus_mlp_scratch_floats(cfg, N), a)
        L.check(lib.fneus_mlp_bwd(cfg, L.ptr(w), L.ptr(a), L.ptr(b) if ctx.has_b else None, N, L.ptr(out),
                                  L.ptr(_f32c(d_out)), L.ptr(saved), L.ptr(scratch), L.ptr(dw), L.stream_ptr()),
                "fneus_mlp_bwd")
        return dw, None, None, None


class NerfMLP(torch.autograd.Function):
    """NeRF.forward with view directions (fields.py:233-259) -> (raw density [N,1], raw rgb [N,3])."""

    @staticmethod
    def forward(ctx, wflat, pts, views, cfg):
        _need_cuda(pts, "pts")
        w, p, v = _f32c(wflat), _f32c(pts), _f32c(views)
        N = p.shape[0]
        lib = L.lib()
        dens = torch.empty(N, 1, dtype=torch.float32, device=p.device)
        rgb = torch.empty(N, 3, dtype=torch.float32, device=p.device)
        saved = _empty(lib.fneus_nerf_saved_floats(cfg, N), p)
        L.check(lib.fneus_nerf_fwd(cfg, L.ptr(w), L.ptr(p), L.ptr(v), N, L.ptr(dens), L.ptr(rgb), L.ptr(saved),
                                   L.stream_ptr()), "fneus_nerf_fwd")
        ctx.cfg = cfg
        ctx.save_for_backward(w, p, v, saved)
        return dens, rgb

    pre_backward_hook = None       # the NeRF backward starts after the SDF backward: the SDF slice is final by then

    @staticmethod
    def backward(ctx, d_dens, d_rgb):
        if NerfMLP.pre_backward_hook is not None:
            NerfMLP.pre_backward_hook()
        w, p, v, saved = ctx.saved_tensors
        cfg = ctx.cfg
        N = p.shape[0]
        lib = L.lib()
        dw = torch.zeros_like(w)
        scratch = _empty(lib.fneus_nerf_scratch_floats(cfg, N), p)
        L.check(lib.fneus_nerf_bwd(cfg, L.ptr(w), L.ptr(p), L.ptr(v), N, L.ptr(_f32c(d_dens)), L.ptr(_f32c(d_rgb)),
                                   L.ptr(saved), L.ptr(scratch), L.ptr(dw), L.stream_ptr()), "fneus_nerf_bwd")
        return dw, None, None, None


class OutsideAlpha(torch.autograd.Function):
    """renderer.py:131-134: (density [N], rgb_raw [N,3], dists [N]) -> (alpha [N], color [N,3])."""

    @staticmethod
    def forward(ctx, density, rgb_raw, dists):
        dn, rg, ds = _f32c(density).reshape(-1), _f32c(rgb_raw).reshape(-1, 3), _f32c(dists).reshape(-1)
        N = dn.shape[0]
        alpha = torch.empty_like(dn)
        color = torch.empty_like(rg)
        L.check(L.lib().fneus_outside_alpha_fwd(L.ptr(dn), L.ptr(rg), L.ptr(ds), N, L.ptr(alpha), L.ptr(color),
                                                L.stream_ptr()), "fneus_outside_alpha_fwd")
        ctx.save_for_backward(dn, color, ds)
        ctx.shapes = (density.shape, rgb_raw.shape)
        return alpha, color

    @staticmethod
    def backward(ctx, d_alpha, d_color):
        dn, color, ds = ctx.saved_tensors
        N = dn.shape[0]
        d_dn = torch.empty_like(dn)
        d_rg = torch.empty_like(color)
        L.check(L.lib().fneus_outside_alpha_bwd(L.ptr(dn), L.ptr(color), L.ptr(ds), L.ptr(_f32c(d_alpha).reshape(-1)),
                                                L.ptr(_f32c(d_color).reshape(-1, 3)), N, L.ptr(d_dn), L.ptr(d_rg),
                                                L.stream_ptr()), "fneus_outside_alpha_bwd")
        return d_dn.reshape(ctx.shapes[0]), d_rg.reshape(ctx.shapes[1]), None


def outside_geometry(rays_o, rays_d, z, sample_dist):
    B, n = z.shape
    dev = z.device
    dists = torch.empty(B, n, dtype=torch.float32, device=dev)
    pts4 = torch.empty(B * n, 4, dtype=torch.float32, device=dev)
    dirs = torch.empty(B * n, 3, dtype=torch.float32, device=dev)
    L.check(L.lib().fneus_outside_geometry(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), B, n, float(sample_dist),
                                           L.ptr(dists), L.ptr(pts4), L.ptr(dirs), L.stream_ptr()),
            "fneus_outside_geometry")
    return dists, pts4, dirs


# ---------------------------------------------------------------------------------------------
# sampling (no grad)
# ---------------------------------------------------------------------------------------------
def ray_points(rays_o, rays_d, z):
    B, n = z.shape
    pts = torch.empty(B * n, 3, dtype=torch.float32, device=z.device)
    L.check(L.lib().fneus_ray_points(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), B, n, L.ptr(pts), L.stream_ptr()),
            "fneus_ray_points")
    return pts


def upsample_step(rays_o, rays_d, z, sdf, k, inv_s, u_table, debug=False):
    B, n = z.shape
    new_z = torch.empty(B, k, dtype=torch.float32, device=z.device)
    cdf = torch.empty(B, n, dtype=torch.float32, device=z.device) if debug else None
    inds = torch.empty(B, k, dtype=torch.int64, device=z.device) if debug else None
    L.check(L.lib().fneus_upsample_step(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(sdf), B, n, k, float(inv_s),
                                        L.ptr(u_table), L.ptr(new_z), L.ptr(cdf), L.ptr(inds), L.stream_ptr()),
            "fneus_upsample_step")
    return (new_z, cdf, inds) if debug else new_z


def upsample_step_dev(rays_o, rays_d, z, sdf, k, inv_s_dev, u_table):
    """up_sample with inv_s taken from a device scalar (the learned inv_s of stage 2; no host read-back)."""
    B, n = z.shape
    new_z = torch.empty(B, k, dtype=torch.float32, device=z.device)
    L.check(L.lib().fneus_upsample_step_dev(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), L.ptr(sdf), B, n, k,
                                            L.ptr(_f32c(inv_s_dev).reshape(-1)[:1].contiguous()), L.ptr(u_table),
                                            L.ptr(new_z), L.stream_ptr()), "fneus_upsample_step_dev")
    return new_z


def upsample_iter(rays_o, rays_d, z, sdf, prev_z, prev_sdf, k, inv_s, u_table, want_pts=True):
    """One iteration of the hierarchical sampling loop (renderer.py:166-176) in one launch: merge (prev_z, prev_sdf) into
    (z, sdf) when given, up-sample k new depths from the merged row and return their positions for the next SDF pass.
    -> (z_merged, sdf_merged, new_z [B,k], pts [B*k,3] or None).  Bit-identical to merge_sorted + upsample_step + ray_points."""
    _need_cuda(z, "z")
    B, n = z.shape
    kp = 0 if prev_z is None else prev_z.shape[1]
    f32 = dict(dtype=torch.float32, device=z.device)
    zc, sc = _f32c(z), _f32c(sdf).reshape(B, n)
    if kp:
        z_out, sdf_out = torch.empty(B, n + kp, **f32), torch.empty(B, n + kp, **f32)
        pz, ps = _f32c(prev_z), _f32c(prev_sdf).reshape(B, kp)
    else:
        z_out, sdf_out, pz, ps = zc, sc, None, None
    new_z = torch.empty(B, k, **f32)
    pts = torch.empty(B * k, 3, **f32) if want_pts else None
    L.check(L.lib().fneus_upsample_iter(L.ptr(_f32c(rays_o)), L.ptr(_f32c(rays_d)), L.ptr(zc), L.ptr(sc), B, n, L.ptr(pz),
                                        L.ptr(ps), kp, int(k), float(inv_s), L.ptr(_f32c(u_table)),
                                        L.ptr(z_out) if kp else None, L.ptr(sdf_out) if kp else None, L.ptr(new_z),
                                        L.ptr(pts), L.stream_ptr()), "fneus_upsample_iter")
    return z_out, sdf_out, new_z, pts


def first_hit_secant(sdf, mid_z, pts, rays_o, rays_d, weights=None):
    """renderer.py:588-602 / calLvis.py:180-196 with fixed shapes: (hit_idx [B] int32, -1 = no hit; z_surf [B,1];
    pts_surf [B,3]; lvis [B] = 1 - sum w * inside when ``weights`` is given, else None; any_inside [B] bool)."""
    _need_cuda(sdf, "sdf")
    B, n = mid_z.shape
    dev = mid_z.device
    sdf_c, mid_c, pts_c = _f32c(sdf).reshape(B, n), _f32c(mid_z), _f32c(pts).reshape(B * n, 3)
    hit = torch.empty(B, dtype=torch.int32, device=dev)
    zs = torch.empty(B, 1, dtype=torch.float32, device=dev)
    ps = torch.empty(B, 3, dtype=torch.float32, device=dev)
    w = _f32c(weights)
    lv = torch.empty(B, dtype=torch.float32, device=dev) if w is not None else None
    anyin = torch.empty(B, dtype=torch.int32, device=dev)
    L.check(L.lib().fneus_first_hit_secant(L.ptr(sdf_c), L.ptr(mid_c), L.ptr(pts_c), L.ptr(_f32c(rays_o)),
                                           L.ptr(_f32c(rays_d)), L.ptr(w), w.shape[1] if w is not None else 0, B, n,
                                           L.ptr(hit), L.ptr(zs), L.ptr(ps), L.ptr(lv), L.ptr(anyin), L.stream_ptr()),
            "fneus_first_hit_secant")
    return hit, zs, ps, lv, anyin != 0


_LVIS_WS = {}


def lvis_trace(sdf_cfg, sdf_w, color_cfg, color_w, surf, dirs, inv_s, z_table, u_table, rays_per_chunk=8192):
    """calLvis.py:351-397 (ground truth of cal_indiLgt) in ONE library call: surf [m,3], dirs [m,k,3] ->
    (gt_lvis [m,k], gt_trace_radiance [m,k,3], hit [m,k] int32).  The workspace is cached per (device, shape)."""
    _need_cuda(surf, "surf")
    s, d = _f32c(surf), _f32c(dirs)
    m, k = d.shape[0], d.shape[1]
    n_coarse, n_imp = z_table.numel(), u_table.numel()
    dev = s.device
    lib = L.lib()
    chunk = max(k, min(rays_per_chunk, m * k) // k * k)
    need = lib.fneus_lvis_trace_workspace_floats(sdf_cfg, color_cfg, chunk, n_coarse, n_imp)
    if need < 0:
        raise RuntimeError("fneus_lvis_trace: unsupported configuration")
    key = (str(dev), chunk, n_coarse, n_imp, int(sdf_cfg.precision))
    ws = _LVIS_WS.get(key)
    if ws is None or ws.numel() < need:
        _LVIS_WS.clear()
        ws = _LVIS_WS[key] = torch.empty(int(need), dtype=torch.float32, device=dev)
    lvis = torch.empty(m, k, dtype=torch.float32, device=dev)
    rgb = torch.empty(m, k, 3, dtype=torch.float32, device=dev)
    hit = torch.empty(m, k, dtype=torch.int32, device=dev)
    L.check(lib.fneus_lvis_trace(sdf_cfg, L.ptr(_f32c(sdf_w)), color_cfg, L.ptr(_f32c(color_w)), L.ptr(s), L.ptr(d.reshape(-1, 3)),
                                 m, k, n_coarse, n_imp, L.ptr(_f32c(inv_s).reshape(-1)[:1].contiguous()), L.ptr(z_table),
                                 L.ptr(u_table), L.ptr(lvis), L.ptr(rgb), L.ptr(hit), L.ptr(ws), ws.numel(), chunk,
                                 L.stream_ptr()), "fneus_lvis_trace")
    return lvis, rgb, hit


def inverse_cdf(bins, cdf, u_table):
    B, n = bins.shape
    k = u_table.numel()
    out = torch.empty(B, k, dtype=torch.float32, device=bins.device)
    inds = torch.empty(B, k, dtype=torch.int64, device=bins.device)
    L.check(L.lib().fneus_inverse_cdf(L.ptr(bins), L.ptr(cdf), L.ptr(u_table), B, n, k, L.ptr(out), L.ptr(inds),
                                      L.stream_ptr()), "fneus_inverse_cdf")
    return out, inds


def merge_sorted(z, new_z, sdf=None, new_sdf=None):
    B, n = z.shape
    k = new_z.shape[1]
    z_out = torch.empty(B, n + k, dtype=torch.float32, device=z.device)
    sdf_out = torch.empty(B, n + k, dtype=torch.float32, device=z.device) if sdf is not None else None
    L.check(L.lib().fneus_merge_sorted(L.ptr(z), L.ptr(new_z), L.ptr(sdf), L.ptr(new_sdf), B, n, k, L.ptr(z_out),
                                       L.ptr(sdf_out), L.stream_ptr()), "fneus_merge_sorted")
    return z_out, sdf_out


def core_geometry(rays_o, rays_d, z, sample_dist):
    B, n = z.shape
    dev = z.device
    dists = torch.empty(B, n, dtype=torch.float32, device=dev)
    mid_z = torch.empty(B, n, dtype=torch.float32, device=dev)
    pts = torch.empty(B * n, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(B * n, 3, dtype=torch.float32, device=dev)
    L.check(L.lib().fneus_core_geometry(L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), B, n, float(sample_dist),
                                        L.ptr(dists), L.ptr(mid_z), L.ptr(pts), L.ptr(dirs), L.stream_ptr()),
            "fneus_core_geometry")
    return dists, mid_z, pts, dirs


# ---------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------
_ONES = {}


def _ones1(dev):
    """A cached [1] tensor of ones per device (never written)."""
    t = _ONES.get(dev)
    if t is None:
        t = _ONES[dev] = torch.ones(1, dtype=torch.float32, device=dev)
    return t


def split_batch(batch):
    """[B,10] rows of Dataset.gen_random_rays_at (dataset.py:133-151) -> rays_o, rays_d, rgb [B,3], mask [B,1] as dense
    tensors in one launch (slicing the columns costs one strided copy per consumer)."""
    _need_cuda(batch, "batch")
    bc = _f32c(batch)
    B = bc.shape[0]
    f32 = dict(dtype=torch.float32, device=bc.device)
    ro, rd, rgb, m = torch.empty(B, 3, **f32), torch.empty(B, 3, **f32), torch.empty(B, 3, **f32), torch.empty(B, 1, **f32)
    L.check(L.lib().fneus_split_batch(L.ptr(bc), B, L.ptr(ro), L.ptr(rd), L.ptr(rgb), L.ptr(m), L.stream_ptr()),
            "fneus_split_batch")
    return ro, rd, rgb, m


class InvS(torch.autograd.Function):
    """variance (scalar parameter) -> inv_s [1,1] = clip(exp(10 variance), 1e-6, 1e6): SingleVarianceNetwork.forward on one
    row (fields.py:267-268) with the clip of renderer.py:238, one launch each way."""

    @staticmethod
    def forward(ctx, variance):
        _need_cuda(variance, "variance")
        v = _f32c(variance).reshape(1)
        out = torch.empty(1, 1, dtype=torch.float32, device=v.device)
        L.check(L.lib().fneus_inv_s(L.ptr(v), None, L.ptr(out), L.stream_ptr()), "fneus_inv_s")
        ctx.save_for_backward(v)
        ctx.shape = variance.shape
        return out

    @staticmethod
    def backward(ctx, d_inv_s):
        (v,) = ctx.saved_tensors
        d = torch.empty(1, dtype=torch.float32, device=v.device)
        L.check(L.lib().fneus_inv_s(L.ptr(v), L.ptr(_f32c(d_inv_s).reshape(1)), L.ptr(d), L.stream_ptr()), "fneus_inv_s")
        return d.reshape(ctx.shape)


class GatherRows3(torch.autograd.Function):
    """(pts, dirs, normals)[rows] in one launch (renderer.py:296-327 gathers the two samples around the sign change);
    only the normals carry a gradient."""

    @staticmethod
    def forward(ctx, pts, dirs, normals, rows):
        _need_cuda(pts, "pts")
        a, b, c = _f32c(pts).reshape(-1, 3), _f32c(dirs).reshape(-1, 3), _f32c(normals).reshape(-1, 3)
        n = rows.shape[0]
        outs = [torch.empty(n, 3, dtype=torch.float32, device=a.device) for _ in range(3)]
        L.check(L.lib().fneus_gather_rows3(L.ptr(a), L.ptr(b), L.ptr(c), L.ptr(rows), n, L.ptr(outs[0]), L.ptr(outs[1]),
                                           L.ptr(outs[2]), L.stream_ptr()), "fneus_gather_rows3")
        ctx.save_for_backward(rows)
        ctx.nshape = normals.shape
        ctx.mark_non_differentiable(outs[0], outs[1])
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, _gp, _gd, g_n):
        if g_n is None:
            return None, None, None, None
        (rows,) = ctx.saved_tensors
        g = torch.zeros(ctx.nshape, dtype=torch.float32, device=g_n.device)
        L.check(L.lib().fneus_scatter_rows3(L.ptr(_f32c(g_n)), L.ptr(rows), rows.shape[0], L.ptr(g), L.stream_ptr()),
                "fneus_scatter_rows3")
        return None, None, g, None


class Composite(torch.autograd.Function):
    """renderer.py:245-274,328-332,350-372 fused; differentiable in sdf, normals, rgb, inv_s, bg_alpha, bg_color.

    Returns (color [B,3], weights [B,n_tot], weight_sum [B,1], weight_max [B,1], cdf [B,n_in],
    inside [B,n_in], eik_num [], eik_den [], hit_idx [B] int32, w_pair [B,2]) where the eikonal term of
    renderer.py:370-372 is eik_num / (eik_den + 1e-5) (kept apart so that ray-sharded data parallelism can
    normalise by the batch-global denominator)."""

    @staticmethod
    def forward(ctx, sdf, normals, rgb, inv_s, bg_alpha, bg_color, dists, pts, rays_d, bg_rgb, n_in, n_out, car):
        _need_cuda(sdf, "sdf")
        dev = sdf.device
        sdf_c, nrm_c, rgb_c = _f32c(sdf).reshape(-1), _f32c(normals).reshape(-1, 3), _f32c(rgb).reshape(-1, 3)
        inv_c = _f32c(inv_s).reshape(-1)[:1].contiguous()
        bga, bgc = _f32c(bg_alpha), _f32c(bg_color)
        dists_c, pts_c, rd_c, bgr = _f32c(dists), _f32c(pts), _f32c(rays_d), _f32c(bg_rgb)
        if bgr is not None:
            bgr = bgr.reshape(-1)[:3].contiguous()
        B = rd_c.shape[0]
        n_tot = n_in + n_out
        # cos_anneal_ratio: a python float, or a DEVICE scalar tensor (read by the kernels: graph-capturable schedule)
        car_dev = _f32c(car).reshape(-1)[:1].contiguous() if torch.is_tensor(car) else None
        car = 0.0 if car_dev is not None else float(car)
        f32 = dict(dtype=torch.float32, device=dev)
        color = torch.empty(B, 3, **f32)
        weights = torch.empty(B, n_tot, **f32)
        wsum = torch.empty(B, 1, **f32)
        wmax = torch.empty(B, 1, **f32)
        cdf = torch.empty(B, n_in, **f32)
        inside = torch.empty(B, n_in, **f32)
        eik = torch.empty(B, 2, **f32)
        hit = torch.empty(B, dtype=torch.int32, device=dev)
        wpair = torch.empty(B, 2, **f32)
        L.check(L.lib().fneus_composite_fwd(
            L.ptr(sdf_c), L.ptr(nrm_c), L.ptr(rgb_c), L.ptr(dists_c), L.ptr(pts_c), L.ptr(rd_c), L.ptr(bga),
            L.ptr(bgc), L.ptr(bgr), B, n_in, n_out, L.ptr(inv_c), car, L.ptr(car_dev), L.ptr(color), L.ptr(weights),
            L.ptr(wsum), L.ptr(wmax), L.ptr(cdf), L.ptr(inside), L.ptr(eik), L.ptr(hit), L.ptr(wpair),
            L.stream_ptr()), "fneus_composite_fwd")
        # ray totals of the eikonal term, their quotient and the sign-change mask: one launch (fixed summation order)
        tot = torch.empty(3, **f32)
        hit_mask = torch.empty(B, dtype=torch.bool, device=dev)
        L.check(L.lib().fneus_composite_post(L.ptr(eik), L.ptr(hit), B, L.ptr(tot), L.ptr(hit_mask), L.stream_ptr()),
                "fneus_composite_post")
        eik_num, eik_den, grad_err = tot[0].reshape(()), tot[1].reshape(()), tot[2].reshape(())
        denom = _ones1(dev)
        ctx.save_for_backward(sdf_c, nrm_c, rgb_c, inv_c, bga, bgc, dists_c, pts_c, rd_c, bgr, hit, denom, tot)
        ctx.dims = (B, n_in, n_out, car)
        ctx.car_dev = car_dev
        ctx.shapes = (sdf.shape, normals.shape, rgb.shape, inv_s.shape)
        ctx.mark_non_differentiable(wmax, cdf, inside, hit, eik_den, hit_mask)
        ctx.set_materialize_grads(False)
        return color, weights, wsum, wmax, cdf, inside, eik_num, eik_den, hit, wpair, grad_err, hit_mask

    @staticmethod
    def backward(ctx, d_color, d_weights, d_wsum, _wmax, _cdf, _inside, d_eik, _eik_den, _hit, d_wpair, d_gerr, _hm):
        sdf_c, nrm_c, rgb_c, inv_c, bga, bgc, dists_c, pts_c, rd_c, bgr, hit, denom, tot = ctx.saved_tensors
        if d_gerr is not None:
            # gradient_error = eik_num / (eik_den + 1e-5) (renderer.py:282) used directly in a loss: fold it into the
            # numerator's gradient (the denominator is not differentiable, as in the reference: relax_inside_sphere is a mask)
            extra = d_gerr.reshape(()) / (tot[1] + 1e-5)
            d_eik = extra if d_eik is None else d_eik.reshape(()) + extra
        B, n_in, n_out, car = ctx.dims
        d_sdf = torch.empty_like(sdf_c)
        d_nrm = torch.empty_like(nrm_c)
        d_rgb = torch.empty_like(rgb_c)
        d_inv = torch.empty(B, dtype=torch.float32, device=sdf_c.device)
        d_bga = torch.empty_like(bga) if bga is not None else None
        d_bgc = torch.empty_like(bgc) if bgc is not None else None
        d_eik_c = _f32c(d_eik).reshape(1) if d_eik is not None else None
        L.check(L.lib().fneus_composite_bwd(
            L.ptr(sdf_c), L.ptr(nrm_c), L.ptr(rgb_c), L.ptr(dists_c), L.ptr(pts_c), L.ptr(rd_c), L.ptr(bga),
            L.ptr(bgc), L.ptr(bgr), B, n_in, n_out, L.ptr(inv_c), car, L.ptr(ctx.car_dev), L.ptr(hit), L.ptr(_f32c(d_color)),
            L.ptr(_f32c(d_weights)), L.ptr(_f32c(d_wsum)), L.ptr(_f32c(d_wpair)), L.ptr(d_eik_c), L.ptr(denom),
            L.ptr(d_sdf), L.ptr(d_nrm), L.ptr(d_rgb), L.ptr(d_inv), L.ptr(d_bga), L.ptr(d_bgc), L.stream_ptr()),
            "fneus_composite_bwd")
        s_sdf, s_nrm, s_rgb, s_inv = ctx.shapes
        d_inv_s = d_inv.sum().reshape(s_inv)
        return (d_sdf.reshape(s_sdf), d_nrm.reshape(s_nrm), d_rgb.reshape(s_rgb), d_inv_s, d_bga, d_bgc,
                None, None, None, None, None, None, None)


class SurfaceBlend(torch.autograd.Function):
    """renderer.py:328-343 for the three RefColor outputs at once: c_* [2B,3] (rows 2b, 2b+1), w_pair [B,2],
    hit_idx [B] int32 -> three [B,3] colours (ones where the ray has no sign change)."""

    @staticmethod
    def forward(ctx, c_rgb, c_spec, c_diff, w_pair, hit_idx):
        _need_cuda(c_rgb, "c_rgb")
        cr, cs, cd, w = _f32c(c_rgb), _f32c(c_spec), _f32c(c_diff), _f32c(w_pair)
        B = w.shape[0]
        outs = [torch.empty(B, 3, dtype=torch.float32, device=w.device) for _ in range(3)]
        L.check(L.lib().fneus_surface_blend_fwd(L.ptr(cr), L.ptr(cs), L.ptr(cd), L.ptr(w), L.ptr(hit_idx), B,
                                                L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(outs[2]), L.stream_ptr()),
                "fneus_surface_blend_fwd")
        ctx.save_for_backward(cr, cs, cd, w, hit_idx)
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_rgb, g_spec, g_diff):
        cr, cs, cd, w, hit_idx = ctx.saved_tensors
        B = w.shape[0]
        ds = [torch.empty_like(cr), torch.empty_like(cs), torch.empty_like(cd)]
        dw = torch.empty_like(w)
        L.check(L.lib().fneus_surface_blend_bwd(L.ptr(cr), L.ptr(cs), L.ptr(cd), L.ptr(w), L.ptr(hit_idx), B,
                                                L.ptr(_f32c(g_rgb)), L.ptr(_f32c(g_spec)), L.ptr(_f32c(g_diff)),
                                                L.ptr(ds[0]), L.ptr(ds[1]), L.ptr(ds[2]), L.ptr(dw), L.stream_ptr()),
                "fneus_surface_blend_bwd")
        return ds[0], ds[1], ds[2], dw, None


def loss_norms(mask, hit_idx, eik_den, use_mask):
    """den4 = [sum mask, sum mask*hit, eik_den, n_rays] (exp_runner.py:146,160; renderer.py:372)."""
    _need_cuda(mask, "mask")
    m = _f32c(mask).reshape(-1)
    den = torch.empty(4, dtype=torch.float32, device=m.device)
    L.check(L.lib().fneus_loss_norms(L.ptr(m), L.ptr(hit_idx), L.ptr(_f32c(eik_den).reshape(1)), m.shape[0],
                                     1 if use_mask else 0, L.ptr(den), L.stream_ptr()), "fneus_loss_norms")
    return den


class Stage1Loss(torch.autograd.Function):
    """exp_runner.py:134-177 in one launch: returns parts5 = [loss, color, surface, eikonal, mask]; gradients flow to
    color_fine, surface_color, weight_sum and eik_num through parts5[0] only."""

    @staticmethod
    def forward(ctx, color, surf, wsum, eik_num, true_rgb, mask, hit_idx, den4, use_mask, sw, iw, mw):
        _need_cuda(color, "color_fine")
        c, s_, w = _f32c(color), _f32c(surf), _f32c(wsum).reshape(-1)
        e = _f32c(eik_num).reshape(1)
        t, m = _f32c(true_rgb), _f32c(mask).reshape(-1)
        B = c.shape[0]
        parts = torch.empty(5, dtype=torch.float32, device=c.device)
        # the four gradients live in one buffer so that the backward scales them with one launch
        gbuf = torch.empty(7 * B + 1, dtype=torch.float32, device=c.device)
        dc, ds, dw, de = gbuf[:3 * B].view(B, 3), gbuf[3 * B:6 * B].view(B, 3), gbuf[6 * B:7 * B], gbuf[7 * B:]
        L.check(L.lib().fneus_stage1_loss(L.ptr(c), L.ptr(s_), L.ptr(w), L.ptr(t), L.ptr(m), L.ptr(hit_idx), L.ptr(e),
                                          L.ptr(den4), B, 1 if use_mask else 0, float(sw), float(iw), float(mw),
                                          L.ptr(parts), L.ptr(dc), L.ptr(ds), L.ptr(dw), L.ptr(de), L.stream_ptr()),
                "fneus_stage1_loss")
        ctx.save_for_backward(gbuf)
        ctx.shapes = (color.shape, surf.shape, wsum.shape, eik_num.shape)
        return parts

    @staticmethod
    def backward(ctx, g_parts):
        (gbuf,) = ctx.saved_tensors
        B = (gbuf.shape[0] - 1) // 7
        g = gbuf * g_parts[0]
        sc, ss, sw_, se = ctx.shapes
        return (g[:3 * B].reshape(sc), g[3 * B:6 * B].reshape(ss), g[6 * B:7 * B].reshape(sw_), g[7 * B:].reshape(se),
                None, None, None, None, None, None, None, None)


class Stage2Loss(torch.autograd.Function):
    """lvis.py:163-170 in one launch: parts3 = [loss, lvis_loss, radiance_loss]; gradients flow to pre_lvis and
    pre_trace_radiance through parts3[0]."""

    @staticmethod
    def forward(ctx, pre_lvis, pre_rad, gt_lvis, gt_rad, hit_idx, den2):
        _need_cuda(pre_lvis, "pre_lvis")
        pl, pr, gl, gr = _f32c(pre_lvis), _f32c(pre_rad), _f32c(gt_lvis), _f32c(gt_rad)
        B, k = pl.shape
        parts = torch.empty(3, dtype=torch.float32, device=pl.device)
        dl, dr = torch.empty_like(pl), torch.empty_like(pr)
        L.check(L.lib().fneus_stage2_loss(L.ptr(gl), L.ptr(pl), L.ptr(gr), L.ptr(pr), L.ptr(hit_idx), L.ptr(_f32c(den2)), B, k,
                                          L.ptr(parts), L.ptr(dl), L.ptr(dr), L.stream_ptr()), "fneus_stage2_loss")
        ctx.save_for_backward(dl, dr)
        return parts

    @staticmethod
    def backward(ctx, g_parts):
        dl, dr = ctx.saved_tensors
        g = g_parts[0]
        return dl * g, dr * g, None, None, None, None


def stage2_loss(out, den2=None):
    """Loss of the stage-2 Runner (lvis.py:163-170) from a ``lvis_render`` output dict: (loss, stats)."""
    hit_idx = torch.where(out["sdf_mask"], 1, -1).to(torch.int32)
    parts = Stage2Loss.apply(out["pre_lvis"], out["pre_trace_radiance"], out["gt_lvis"], out["gt_trace_radiance"], hit_idx,
                             den2)
    st = parts.detach()
    return parts[0], dict(lvis_loss=st[1], radiance_loss=st[2])


def gen_rays(px, py, intrinsics_inv, pose, image=None, mask=None, with_near_far=True):
    """dataset.py:115-151 on the device: [B,10] = (rays_o, rays_v, rgb, mask) from pixel coordinates, plus near / far
    (dataset.py:186-192).  intrinsics_inv, pose: 4x4; image, mask: [H,W,3] device tensors or None."""
    _need_cuda(px, "pixel coordinates")
    x, y = _f32c(px).reshape(-1), _f32c(py).reshape(-1)
    B = x.shape[0]
    dev = x.device
    out = torch.empty(B, 10, dtype=torch.float32, device=dev)
    near = torch.empty(B, 1, dtype=torch.float32, device=dev) if with_near_far else None
    far = torch.empty(B, 1, dtype=torch.float32, device=dev) if with_near_far else None
    H, W = (image.shape[0], image.shape[1]) if image is not None else ((mask.shape[0], mask.shape[1]) if mask is not None else (0, 0))
    L.check(L.lib().fneus_gen_rays(L.ptr(x), L.ptr(y), L.ptr(_f32c(intrinsics_inv)), L.ptr(_f32c(pose)), L.ptr(_f32c(image)),
                                   L.ptr(_f32c(mask)), H, W, B, L.ptr(out), L.ptr(near), L.ptr(far), L.stream_ptr()),
            "fneus_gen_rays")
    return out, near, far


def adam_step(p, g, m, v, state4, base_lr, lr_alpha, warm_up_end, end_iter, beta1=0.9, beta2=0.999, eps=1e-8,
              grad_scale=1.0, zero_grad=True):
    """Fused flat Adam with the on-device warm-up/cosine schedule (exp_runner.py:118,229-238); in place."""
    _need_cuda(p, "params")
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
            raise RuntimeError("adam_step: flat contiguous FP32 buffers of equal length expected")
    L.check(L.lib().fneus_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), L.ptr(state4), float(base_lr),
                                    float(lr_alpha), float(warm_up_end), float(end_iter), float(beta1), float(beta2),
                                    float(eps), float(grad_scale), 1 if zero_grad else 0, L.stream_ptr()),
            "fneus_adam_step")


def near_far_from_sphere(rays_o, rays_d):
    """dataset.py:186-192 in one launch: (near, far) [B,1]."""
    _need_cuda(rays_o, "rays_o")
    ro, rd = _f32c(rays_o), _f32c(rays_d)
    B = ro.shape[0]
    near = torch.empty(B, 1, dtype=torch.float32, device=ro.device)
    far = torch.empty(B, 1, dtype=torch.float32, device=ro.device)
    L.check(L.lib().fneus_near_far(L.ptr(ro), L.ptr(rd), B, L.ptr(near), L.ptr(far), L.stream_ptr()), "fneus_near_far")
    return near, far


def coarse_z(near, far, lin, rnd, n_samples):
    """renderer.py:395-408: z [B,n] = near + (far - near) * lin (+ (rnd - 0.5) * 2 / n_samples)."""
    _need_cuda(near, "near")
    nr, fr = _f32c(near).reshape(-1), _f32c(far).reshape(-1)
    B, n = nr.shape[0], lin.numel()
    z = torch.empty(B, n, dtype=torch.float32, device=nr.device)
    r = _f32c(rnd).reshape(-1) if rnd is not None else None
    inv_n = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(n_samples), dtype=torch.float32))
    L.check(L.lib().fneus_coarse_z(L.ptr(nr), L.ptr(fr), L.ptr(lin), L.ptr(r), B, n, inv_n, L.ptr(z),
                                   L.stream_ptr()), "fneus_coarse_z")
    return z


def hit_rows(hit_idx, n):
    """renderer.py:296-303: flat rows (idx-1, idx) per ray, int64 [2B]."""
    B = hit_idx.shape[0]
    rows = torch.empty(2 * B, dtype=torch.int64, device=hit_idx.device)
    L.check(L.lib().fneus_hit_rows(L.ptr(hit_idx), B, int(n), L.ptr(rows), L.stream_ptr()), "fneus_hit_rows")
    return rows


class FanOut(torch.autograd.Function):
    """Identity with two outputs for a big activation that feeds one DENSE consumer (the colour network reads every row
    of `feature`) and one SPARSE consumer (RefColor reads 2 rows per ray, renderer.py:296-327).  Autograd would
    materialise the sparse consumer's gradient as a dense zero-filled tensor and add two dense tensors (3 x 67 MB of
    traffic at 512 rays); here the sparse part is handed over through `stash` by GatherRows.backward and scattered into
    the dense gradient in place.  The engine runs this node only after BOTH branches have delivered (or skipped) their
    gradient, so the hand-over needs no ordering assumption."""

    @staticmethod
    def forward(ctx, x, stash):
        ctx.stash = stash
        ctx.set_materialize_grads(False)
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g_dense, g_other):
        stash = ctx.stash
        g = g_dense
        if g_other is not None:
            if stash.get("feat_image"):
                # an operand image is an opaque byte pattern: two of them cannot be added as float32 tensors
                raise RuntimeError("fneus FanOut: the sparse branch of a feature IMAGE may only feed ops.GatherRows")
            g = g_other if g is None else g + g_other
        pend = stash.pop("rows_grad", None)
        if pend is not None:
            rows, vals, shape = pend
            if g is None:
                g = torch.zeros(shape, dtype=vals.dtype, device=vals.device)
            elif not g.is_contiguous():
                g = g.contiguous()
            if stash.get("feat_image"):
                # `g` is the colour network's BF16 gradient image: the sparse rows are added into it in place
                L.check(L.lib().fneus_image_scatter_add_rows(L.ptr(g), vals.shape[1], L.ptr(rows), rows.shape[0],
                                                             L.ptr(vals), L.stream_ptr()), "fneus_image_scatter_add_rows")
            else:
                g.index_add_(0, rows, vals)
        return g, None


class GatherRows(torch.autograd.Function):
    """x.index_select(0, rows) whose backward hands (rows, grad) to the FanOut node that produced x instead of
    materialising a dense gradient."""

    @staticmethod
    def forward(ctx, x, rows, stash):
        ctx.save_for_backward(rows)
        ctx.stash, ctx.shape = stash, x.shape
        if stash.get("feat_image"):
            return image_gather_rows(x, rows)
        return x.index_select(0, rows)

    @staticmethod
    def backward(ctx, g):
        (rows,) = ctx.saved_tensors
        prev = ctx.stash.get("rows_grad")
        if prev is not None:                         # a second gather off the same fan-out: concatenate
            rows, g = torch.cat([prev[0], rows]), torch.cat([prev[1], g])
        ctx.stash["rows_grad"] = (rows, g.contiguous(), ctx.shape)
        return None, None, None
