"""ctypes access to the CPU emulator of the CUDA kernels (tests/emu/cdp_emu.cpp).

Test infrastructure: lets the CPU-only test tier run the kernel bodies and launch planning of
codeps_b200/csrc on host memory.  Never imported by the codeps_b200 package.
"""
import ctypes
import os
import subprocess
from ctypes import c_float, c_int32, c_size_t, c_void_p

import numpy as np
import torch

from codeps_b200._native import MAX_LEVELS, PhotoArgs

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "cdp_emu.cpp")
LIB = os.path.join(HERE, "emu", "libcdp_emu.so")
DEPS = [SRC] + [os.path.join(REPO, "codeps_b200", "csrc", n)
                for n in ("cdp_common.h", "cdp_math.h", "cdp_kernels.h", "cdp_photo_tile.h", "cdp_plan.h", "cdp_flow.h", "cdp_c2c.h", "cdp_metrics.h")]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
               "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "codeps_b200", "csrc"),
               "-o", LIB, SRC]
        subprocess.run(cmd, check=True)
    lib = ctypes.CDLL(LIB)
    lib.emu_resize_tables_bytes.restype = c_size_t
    lib.emu_photo_scratch_bytes.restype = c_size_t
    lib.emu_photo_saved_bytes.restype = c_size_t
    lib.emu_smooth_saved_bytes.restype = c_size_t
    for name in ("emu_resize_tables_bytes",):
        getattr(lib, name).argtypes = [c_int32] * 3
    for name in ("emu_photo_scratch_bytes", "emu_photo_saved_bytes"):
        getattr(lib, name).argtypes = [c_int32] * 5
    lib.emu_smooth_saved_bytes.argtypes = [c_int32] * 3
    _lib = lib
    return lib


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(None)


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def tile_table(h, w, levels):
    """(level, x0, y0) of every block of one image as the photo kernel's cdp_tile_ctx computes them."""
    lib = load()
    cap = 1 << 20
    out = np.zeros(3 * cap, dtype=np.int32)
    n = lib.emu_tile_table(h, w, levels, out.ctypes.data_as(ctypes.c_void_p), cap)
    assert n > 0, n
    return out[:3 * n].reshape(n, 3)


def resize_tables(h, w, levels):
    lib = load()
    n = lib.emu_resize_tables_bytes(h, w, levels)
    buf = torch.zeros(n, dtype=torch.uint8)
    assert lib.emu_resize_tables_build(c_int32(h), c_int32(w), c_int32(levels), _p(buf)) == 0
    return buf


def photo(level_intrinsics, images, depth, poses, noise, num_scales, alpha=0.85, with_grad=True,
          grad_loss=1.0, seed=0, motions=None, full_res_intrinsics=None):
    """Run emu_photo_fwd (+ emu_photo_bwd).  level_intrinsics: [L,B,4] float32, or None with
    full_res_intrinsics [B,4] to exercise the "intrinsics from device memory" instantiation."""
    lib = load()
    b, _, h, w = depth.shape
    tgt, s0, s1 = (_f32(i) for i in images)
    depth, p0, p1 = _f32(depth), _f32(poses[0]), _f32(poses[1])
    if full_res_intrinsics is not None:
        assert level_intrinsics is None
        k = np.ascontiguousarray(full_res_intrinsics, dtype=np.float32)
        assert k.shape == (b, 4)
    else:
        k = np.ascontiguousarray(level_intrinsics, dtype=np.float32)
        assert k.shape == (num_scales, b, 4)
    tables = resize_tables(h, w, num_scales)
    hm = int(motions is not None)
    mo = [_f32(m) for m in motions] if motions is not None else None
    scratch = torch.zeros(lib.emu_photo_scratch_bytes(b, h, w, num_scales, hm), dtype=torch.uint8)
    saved = torch.zeros(lib.emu_photo_saved_bytes(b, h, w, num_scales, hm), dtype=torch.uint8)
    loss = torch.zeros(1)
    argmin = [torch.full((b, h >> s, w >> s), 77, dtype=torch.uint8) for s in range(num_scales)]
    noise = [_f32(n) for n in noise] if noise is not None else None
    a = PhotoArgs()
    a.batch, a.height, a.width, a.num_levels = b, h, w, num_scales
    a.alpha, a.with_grad = alpha, int(with_grad)
    if full_res_intrinsics is not None:
        a.intrinsics_dev = k.ctypes.data
    else:
        a.intrinsics_host = k.ctypes.data
    a.target, a.source0, a.source1, a.depth = tgt.data_ptr(), s0.data_ptr(), s1.data_ptr(), depth.data_ptr()
    a.pose0, a.pose1 = p0.data_ptr(), p1.data_ptr()
    for s in range(num_scales):
        a.noise[s] = noise[s].data_ptr() if noise is not None else None
        a.argmin[s] = argmin[s].data_ptr()
    a.noise_seed = seed
    a.resize_tables = tables.data_ptr()
    a.loss = loss.data_ptr()
    a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
    a.saved, a.saved_bytes = saved.data_ptr(), saved.numel()
    if mo is not None:
        a.motion0, a.motion1 = mo[0].data_ptr(), mo[1].data_ptr()
    assert lib.emu_photo_fwd(ctypes.byref(a)) == 0
    out = {"recon": loss[0].clone(), "argmin": argmin}
    if with_grad:
        go = torch.tensor([grad_loss], dtype=torch.float32)
        gd = torch.zeros_like(depth)
        g0, g1 = torch.zeros(b, 4, 4), torch.zeros(b, 4, 4)
        gm0 = torch.zeros(b, 3, h, w) if mo is not None else None
        gm1 = torch.zeros(b, 3, h, w) if mo is not None else None
        assert lib.emu_photo_bwd(c_int32(b), c_int32(h), c_int32(w), c_int32(num_scales), _p(saved), _p(tables),
                                 _p(go), _p(gd), _p(g0), _p(g1), c_int32(hm), _p(gm0), _p(gm1)) == 0
        out.update(grad_depth=gd, grad_pose=[g0, g1], grad_motion=[gm0, gm1])
    return out


def photo_from_heads(level_intrinsics, images, disp, pose_params, noise, num_scales, alpha=0.85, grad_loss=1.0,
                     min_depth=0.1, max_depth=100.0):
    """emu_photo_fwd with cdp_photo_args.heads + emu_photo_bwd_heads.  pose_params: ((aa0, t0), (aa1, t1)),
    each [B,3]; the first pair is inverted."""
    from codeps_b200._native import PhotoHeads
    lib = load()
    b, _, h, w = disp.shape
    tgt, s0, s1 = (_f32(i) for i in images)
    disp = _f32(disp)
    flat = [_f32(t.reshape(b, 3)) for pair in pose_params for t in pair]
    k = np.ascontiguousarray(level_intrinsics, dtype=np.float32)
    tables = resize_tables(h, w, num_scales)
    scratch = torch.zeros(lib.emu_photo_scratch_bytes(b, h, w, num_scales, 0), dtype=torch.uint8)
    saved = torch.zeros(lib.emu_photo_saved_bytes(b, h, w, num_scales, 0), dtype=torch.uint8)
    loss = torch.zeros(1)
    depth = torch.zeros(b, 1, h, w)
    p0, p1 = torch.zeros(b, 4, 4), torch.zeros(b, 4, 4)
    argmin = [torch.full((b, h >> s, w >> s), 77, dtype=torch.uint8) for s in range(num_scales)]
    noise = [_f32(n) for n in noise]
    hd = PhotoHeads()
    hd.disp, hd.min_depth, hd.max_depth = disp.data_ptr(), min_depth, max_depth
    hd.axisangle[0], hd.translation[0], hd.axisangle[1], hd.translation[1] = (t.data_ptr() for t in flat)
    hd.invert[0], hd.invert[1] = 1, 0
    a = PhotoArgs()
    a.batch, a.height, a.width, a.num_levels = b, h, w, num_scales
    a.alpha, a.with_grad = alpha, 1
    a.intrinsics_host = k.ctypes.data
    a.target, a.source0, a.source1, a.depth = tgt.data_ptr(), s0.data_ptr(), s1.data_ptr(), depth.data_ptr()
    a.pose0, a.pose1 = p0.data_ptr(), p1.data_ptr()
    for s in range(num_scales):
        a.noise[s] = noise[s].data_ptr()
        a.argmin[s] = argmin[s].data_ptr()
    a.resize_tables = tables.data_ptr()
    a.loss = loss.data_ptr()
    a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
    a.saved, a.saved_bytes = saved.data_ptr(), saved.numel()
    a.heads = ctypes.pointer(hd)
    assert lib.emu_photo_fwd(ctypes.byref(a)) == 0
    go = torch.tensor([grad_loss], dtype=torch.float32)
    gdisp = torch.zeros_like(disp)
    grads = [torch.zeros(b, 3) for _ in range(4)]
    assert lib.emu_photo_bwd_heads(c_int32(b), c_int32(h), c_int32(w), c_int32(num_scales), _p(saved), _p(tables), _p(go),
                                   ctypes.byref(hd), _p(depth), _p(gdisp), *(_p(g) for g in grads)) == 0
    return {"recon": loss[0].clone(), "argmin": argmin, "depth": depth, "poses": [p0, p1], "grad_disp": gdisp,
            "grad_pose_params": grads}


def smooth(image, disp, with_grad=True, grad_loss=1.0):
    lib = load()
    b, _, h, w = disp.shape
    image, disp = _f32(image), _f32(disp)
    saved = torch.zeros(lib.emu_smooth_saved_bytes(b, h, w), dtype=torch.uint8)
    loss = torch.zeros(1)
    assert lib.emu_smooth_fwd(_p(image), _p(disp), c_int32(b), c_int32(h), c_int32(w), c_int32(int(with_grad)),
                              _p(loss), _p(saved)) == 0
    out = {"smooth": loss[0].clone()}
    if with_grad:
        go = torch.tensor([grad_loss], dtype=torch.float32)
        gd = torch.zeros_like(disp)
        assert lib.emu_smooth_bwd(_p(saved), _p(go), c_int32(b), c_int32(h), c_int32(w), _p(gd)) == 0
        out["grad_disp"] = gd
    return out


def flow_loss(maps, kind, wrap_around=True):
    """kind: "smooth" | "sparsity"; returns (loss, [unit gradient per map])."""
    import ctypes
    lib = load()
    maps = [_f32(m) for m in maps]
    n = len(maps)
    b, c, h, w = maps[0].shape
    ptrs = (ctypes.c_void_p * n)(*[m.data_ptr() for m in maps])
    loss = torch.zeros(1)
    unit = torch.zeros(n, b, c, h, w)
    rc = lib.emu_flow_loss(ptrs, c_int32(n), c_int32(b * c), c_int32(h), c_int32(w),
                           c_int32(1 if kind == "sparsity" else 0), c_int32(int(wrap_around)), _p(loss), _p(unit))
    assert rc == 0, rc
    return loss[0].clone(), [unit[i] for i in range(n)]


def depth_metrics(gt, pred, lo, hi, use_gt_scale, garg_crop=False, labels=None, class_id=0):
    """gt / pred [B,1,H,W]; returns the 8 output floats (7 statistics + units with ground truth)."""
    from ctypes import c_int64
    lib = load()
    gt, pred = _f32(gt), _f32(pred)
    b, _, h, w = gt.shape
    units, n = (1, b * h * w) if labels is not None else (b, h * w)
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    out = torch.zeros(8)
    rc = lib.emu_depth_metrics(_p(gt), _p(pred), _p(labels), c_int64(int(class_id)), c_int32(units), c_int32(n),
                               c_int32(h), c_int32(w), c_int32(int(garg_crop)), c_float(lo), c_float(hi),
                               c_int32(int(use_gt_scale)), _p(out))
    assert rc == 0, rc
    return out


def warp_c2c(src, k_src, k_tgt, out_hw, depth_val=1.0, interp_mode="bilinear", padding_mode="border"):
    from ctypes import c_double
    lib = load()
    if src.dim() == 3:
        src = src.unsqueeze(1)
    if src.dtype not in (torch.float32, torch.float64):
        src = src.double()
    src = src.contiguous()
    b, c, hs, ws = src.shape
    ks = np.ascontiguousarray(k_src, dtype=np.float64)
    kt = np.ascontiguousarray(k_tgt, dtype=np.float64)
    out = torch.zeros(b, c, out_hw[0], out_hw[1], dtype=torch.float64)
    rc = lib.emu_warp_c2c(_p(src), c_int32(int(src.dtype == torch.float64)), c_int32(b), c_int32(c), c_int32(hs),
                          c_int32(ws), c_int32(out_hw[0]), c_int32(out_hw[1]), c_void_p(ks.ctypes.data),
                          c_void_p(kt.ctypes.data), c_double(float(depth_val)), c_int32(int(interp_mode == "nearest")),
                          c_int32(int(padding_mode == "zeros")), _p(out))
    assert rc == 0, rc
    return out


def warp_grid(depth, pose, k, motion=None):
    lib = load()
    b, _, h, w = depth.shape
    depth, pose = _f32(depth), _f32(pose)
    motion = _f32(motion) if motion is not None else None
    k = np.ascontiguousarray(k, dtype=np.float32)
    grid = torch.zeros(b, h, w, 2)
    assert lib.emu_warp_grid_fwd(_p(depth), _p(pose), _p(motion), c_void_p(k.ctypes.data), c_int32(b), c_int32(h),
                                 c_int32(w), _p(grid)) == 0
    return grid


def warp_image(src, depth, pose, k, mode=0, motion=None):
    lib = load()
    b, c, h, w = src.shape
    src, depth, pose = _f32(src), _f32(depth), _f32(pose)
    motion = _f32(motion) if motion is not None else None
    k = np.ascontiguousarray(k, dtype=np.float32)
    out = torch.zeros_like(src)
    assert lib.emu_warp_image_fwd(_p(src), c_int32(c), _p(depth), _p(pose), _p(motion), c_void_p(k.ctypes.data),
                                  c_int32(b), c_int32(h), c_int32(w), c_int32(mode), _p(out)) == 0
    return out


def warp_image_bwd(grad_out, src, depth, pose, k, motion=None):
    lib = load()
    b, c, h, w = src.shape
    grad_out, src, depth, pose = _f32(grad_out), _f32(src), _f32(depth), _f32(pose)
    motion = _f32(motion) if motion is not None else None
    k = np.ascontiguousarray(k, dtype=np.float32)
    gd = torch.zeros_like(depth)
    gp = torch.zeros(b, 4, 4)
    gm = torch.zeros_like(motion) if motion is not None else None
    assert lib.emu_warp_image_bwd(_p(grad_out), _p(src), c_int32(c), _p(depth), _p(pose), _p(motion),
                                  c_void_p(k.ctypes.data), c_int32(b), c_int32(h), c_int32(w), _p(gd), _p(gp),
                                  _p(gm)) == 0
    return gd, gp, gm


def ssim(x, y):
    lib = load()
    b, c, h, w = x.shape
    x, y = _f32(x), _f32(y)
    out = torch.zeros_like(x)
    assert lib.emu_ssim_fwd(_p(x), _p(y), c_int32(b * c), c_int32(h), c_int32(w), _p(out)) == 0
    return out


def ssim_bwd(grad_out, x, y):
    lib = load()
    b, c, h, w = x.shape
    grad_out, x, y = _f32(grad_out), _f32(x), _f32(y)
    gx, gy = torch.zeros_like(x), torch.zeros_like(y)
    assert lib.emu_ssim_bwd(_p(grad_out), _p(x), _p(y), c_int32(b * c), c_int32(h), c_int32(w), _p(gx), _p(gy)) == 0
    return gx, gy


def pose(axisangle, translation, invert, grad_out=None):
    lib = load()
    b = axisangle.shape[0]
    aa, tr = _f32(axisangle.reshape(b, 3)), _f32(translation.reshape(b, 3))
    m = torch.zeros(b, 4, 4)
    assert lib.emu_pose_fwd(_p(aa), _p(tr), c_int32(b), c_int32(int(invert)), _p(m)) == 0
    if grad_out is None:
        return m
    go = _f32(grad_out)
    ga, gt = torch.zeros(b, 3), torch.zeros(b, 3)
    assert lib.emu_pose_bwd(_p(go), _p(aa), _p(tr), c_int32(b), c_int32(int(invert)), _p(ga), _p(gt)) == 0
    return m, ga, gt


_ = (MAX_LEVELS, c_float)
