"""torch.autograd glue over the C ABI of libcodeps_photo.so.

Only plumbing lives here: argument checks, buffer allocation through torch's caching allocator,
raw pointers + the current CUDA stream handed to the library, and explicit forward/backward
``autograd.Function``s.  There is deliberately no CPU or eager-PyTorch implementation: tensors
that are not fp32 CUDA tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from ._native import PhotoArgs, check

_LAUNCHES = {"count": 0}  # kernels launched through this module (bench.py reports it)


def launch_count() -> int:
    return _LAUNCHES["count"]


def _require_cuda_f32(t: torch.Tensor, name: str, shape: Optional[Sequence[Optional[int]]] = None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}; codeps_b200 runs on CUDA only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype} (the reference path is fp32 only)")
    if shape is not None:
        if t.dim() != len(shape) or any(s is not None and s != d for s, d in zip(shape, t.shape)):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(None)


def _bytes(n: int, device) -> torch.Tensor:
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


_device_checked = set()


def _lib_for(device):
    lib = _native.load()
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _device_checked:
        with torch.cuda.device(idx):
            check(lib.cdp_device_check(), "cdp_device_check")
        _device_checked.add(idx)
    return lib


def _seed_counter(seed: torch.Tensor, device) -> torch.Tensor:
    if seed.dtype != torch.int64 or seed.numel() != 1 or seed.device != device:
        raise TypeError("a device seed counter must be a one-element int64 tensor on the loss's device")
    return seed


_table_cache = {}


def resize_tables(height: int, width: int, num_levels: int, device) -> torch.Tensor:
    """Device copy of the bilinear-resize tap tables for this pyramid (cached per device)."""
    key = (height, width, num_levels, str(torch.device(device)))
    tab = _table_cache.get(key)
    if tab is None:
        lib = _native.load()
        n = lib.cdp_resize_tables_bytes(height, width, num_levels)
        if n == 0:
            raise ValueError(f"unsupported pyramid: {width}x{height} with {num_levels} levels "
                             f"(at most {_native.MAX_LEVELS} levels, every level at least 2x2)")
        host = torch.empty(n, dtype=torch.uint8)
        check(lib.cdp_resize_tables_build(height, width, num_levels, _ptr(host), n),
              "cdp_resize_tables_build")
        tab = host.to(device)
        _table_cache[key] = tab
    return tab


class PhotoState:
    """Side outputs of one reconstruction-loss call."""
    __slots__ = ("argmin",)

    def __init__(self):
        self.argmin: List[torch.Tensor] = []


class _PhotometricLoss(torch.autograd.Function):
    """cdp_photo_fwd / cdp_photo_bwd."""

    @staticmethod
    def forward(ctx, depth, pose0, pose1, motion0, motion1, target, source0, source1, intrinsics, noise, seed,
                num_levels, alpha, state, noise_event=None):
        b, _, h, w = target.shape
        device = target.device
        lib = _lib_for(device)
        need_grad = any(ctx.needs_input_grad[:5])
        has_motion = motion0 is not None
        tables = resize_tables(h, w, num_levels, device)
        scratch = _bytes(lib.cdp_photo_scratch_bytes(b, h, w, num_levels, int(has_motion)), device)
        saved = _bytes(lib.cdp_photo_saved_bytes(b, h, w, num_levels, int(has_motion)), device) if need_grad else None
        loss = torch.empty(1, dtype=torch.float32, device=device)
        argmin = [torch.empty(b, h >> s, w >> s, dtype=torch.uint8, device=device)
                  for s in range(num_levels)]
        a = PhotoArgs()
        a.batch, a.height, a.width, a.num_levels = b, h, w, num_levels
        a.alpha, a.with_grad = float(alpha), int(need_grad)
        if isinstance(intrinsics, torch.Tensor):
            a.intrinsics_dev = intrinsics.data_ptr()
        else:
            a.intrinsics_host = intrinsics.ctypes.data
        a.target, a.source0, a.source1 = target.data_ptr(), source0.data_ptr(), source1.data_ptr()
        a.depth, a.pose0, a.pose1 = depth.data_ptr(), pose0.data_ptr(), pose1.data_ptr()
        for s in range(num_levels):
            a.noise[s] = noise[s].data_ptr() if noise is not None else None
            a.argmin[s] = argmin[s].data_ptr()
        if isinstance(seed, torch.Tensor):  # device counter: fresh draws on every CUDA-graph replay
            a.noise_seed_dev = _seed_counter(seed, device).data_ptr()
        else:
            a.noise_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        if noise_event is not None:  # the noise was produced on another stream
            a.noise_ready = noise_event.cuda_event
        a.resize_tables = tables.data_ptr()
        a.loss = loss.data_ptr()
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        if need_grad:
            a.saved, a.saved_bytes = saved.data_ptr(), saved.numel()
        if has_motion:
            a.motion0, a.motion1 = motion0.data_ptr(), motion1.data_ptr()
        with torch.cuda.device(device):
            check(lib.cdp_photo_fwd(ctypes.byref(a), _stream(device)), "cdp_photo_fwd")
        _LAUNCHES["count"] += lib.cdp_photo_fwd_launches(b, num_levels)
        state.argmin = argmin
        ctx.shape = (b, h, w, num_levels)
        ctx.saved_buf = saved
        ctx.tables = tables
        ctx.has_motion = has_motion
        return loss[0]

    @staticmethod
    def backward(ctx, grad_loss):
        b, h, w, num_levels = ctx.shape
        saved = ctx.saved_buf
        if saved is None:
            raise RuntimeError("photometric loss: backward called but no input required grad")
        device = saved.device
        lib = _lib_for(device)
        go = _require_cuda_f32(grad_loss.reshape(1), "grad_loss")
        grad_depth = torch.empty(b, 1, h, w, dtype=torch.float32, device=device)
        grad_pose0 = torch.empty(b, 4, 4, dtype=torch.float32, device=device)
        grad_pose1 = torch.empty(b, 4, 4, dtype=torch.float32, device=device)
        gm0 = gm1 = None
        if ctx.has_motion:
            gm0 = torch.empty(b, 3, h, w, dtype=torch.float32, device=device)
            gm1 = torch.empty(b, 3, h, w, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(lib.cdp_photo_bwd(b, h, w, num_levels, _ptr(saved), saved.numel(), _ptr(ctx.tables),
                                    _ptr(go), _ptr(grad_depth), _ptr(grad_pose0), _ptr(grad_pose1),
                                    int(ctx.has_motion), _ptr(gm0), _ptr(gm1), _stream(device)), "cdp_photo_bwd")
        _LAUNCHES["count"] += lib.cdp_photo_bwd_launches(b, num_levels, int(ctx.has_motion))
        return (grad_depth, grad_pose0, grad_pose1, gm0, gm1) + (None,) * 10


class _PhotometricLossHeads(torch.autograd.Function):
    """cdp_photo_fwd with cdp_photo_args.heads / cdp_photo_bwd_heads: the op takes the sigmoid
    disparity and the 6-DoF pose parameters and returns gradients with respect to them; the depth
    map and the two pose matrices come back as (non-differentiable) outputs."""

    @staticmethod
    def forward(ctx, disp, aa0, tr0, aa1, tr1, motion0, motion1, target, source0, source1, intrinsics, noise, seed,
                num_levels, alpha, min_depth, max_depth, state, noise_event=None):
        b, _, h, w = target.shape
        device = target.device
        lib = _lib_for(device)
        need_grad = any(ctx.needs_input_grad[:7])
        has_motion = motion0 is not None
        tables = resize_tables(h, w, num_levels, device)
        scratch = _bytes(lib.cdp_photo_scratch_bytes(b, h, w, num_levels, int(has_motion)), device)
        saved = _bytes(lib.cdp_photo_saved_bytes(b, h, w, num_levels, int(has_motion)), device) if need_grad else None
        loss = torch.empty(1, dtype=torch.float32, device=device)
        depth = torch.empty(b, 1, h, w, dtype=torch.float32, device=device)
        pose0 = torch.empty(b, 4, 4, dtype=torch.float32, device=device)
        pose1 = torch.empty(b, 4, 4, dtype=torch.float32, device=device)
        argmin = [torch.empty(b, h >> s, w >> s, dtype=torch.uint8, device=device) for s in range(num_levels)]
        hd = _native.PhotoHeads()
        hd.disp, hd.min_depth, hd.max_depth = disp.data_ptr(), float(min_depth), float(max_depth)
        hd.axisangle[0], hd.axisangle[1] = aa0.data_ptr(), aa1.data_ptr()
        hd.translation[0], hd.translation[1] = tr0.data_ptr(), tr1.data_ptr()
        hd.invert[0], hd.invert[1] = 1, 0  # t -> t-1 is predicted in temporal order and inverted (algos/depth.py:404-407)
        a = PhotoArgs()
        a.batch, a.height, a.width, a.num_levels = b, h, w, num_levels
        a.alpha, a.with_grad = float(alpha), int(need_grad)
        if isinstance(intrinsics, torch.Tensor):
            a.intrinsics_dev = intrinsics.data_ptr()
        else:
            a.intrinsics_host = intrinsics.ctypes.data
        a.target, a.source0, a.source1 = target.data_ptr(), source0.data_ptr(), source1.data_ptr()
        a.depth, a.pose0, a.pose1 = depth.data_ptr(), pose0.data_ptr(), pose1.data_ptr()
        for s in range(num_levels):
            a.noise[s] = noise[s].data_ptr() if noise is not None else None
            a.argmin[s] = argmin[s].data_ptr()
        if isinstance(seed, torch.Tensor):  # device counter: fresh draws on every CUDA-graph replay
            a.noise_seed_dev = _seed_counter(seed, device).data_ptr()
        else:
            a.noise_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        if noise_event is not None:
            a.noise_ready = noise_event.cuda_event
        a.resize_tables = tables.data_ptr()
        a.loss = loss.data_ptr()
        a.scratch, a.scratch_bytes = scratch.data_ptr(), scratch.numel()
        if need_grad:
            a.saved, a.saved_bytes = saved.data_ptr(), saved.numel()
        if has_motion:
            a.motion0, a.motion1 = motion0.data_ptr(), motion1.data_ptr()
        a.heads = ctypes.pointer(hd)
        with torch.cuda.device(device):
            check(lib.cdp_photo_fwd(ctypes.byref(a), _stream(device)), "cdp_photo_fwd")
        _LAUNCHES["count"] += lib.cdp_photo_fwd_launches(b, num_levels)
        state.argmin = argmin
        ctx.shape = (b, h, w, num_levels)
        ctx.saved_buf = saved
        ctx.tables = tables
        ctx.has_motion = has_motion
        ctx.depth_range = (float(min_depth), float(max_depth))
        ctx.save_for_backward(disp, aa0, tr0, aa1, tr1, depth)
        ctx.mark_non_differentiable(depth, pose0, pose1)
        return loss[0], depth, pose0, pose1

    @staticmethod
    def backward(ctx, grad_loss, _gd, _gp0, _gp1):
        b, h, w, num_levels = ctx.shape
        saved = ctx.saved_buf
        if saved is None:
            raise RuntimeError("photometric loss: backward called but no input required grad")
        disp, aa0, tr0, aa1, tr1, depth = ctx.saved_tensors
        device = saved.device
        lib = _lib_for(device)
        go = _require_cuda_f32(grad_loss.reshape(1), "grad_loss")
        grad_disp = torch.empty_like(disp)
        ga0, gt0, ga1, gt1 = (torch.empty_like(t) for t in (aa0, tr0, aa1, tr1))
        gm0 = gm1 = None
        if ctx.has_motion:
            gm0 = torch.empty(b, 3, h, w, dtype=torch.float32, device=device)
            gm1 = torch.empty(b, 3, h, w, dtype=torch.float32, device=device)
        hd = _native.PhotoHeads()
        hd.disp, hd.min_depth, hd.max_depth = disp.data_ptr(), ctx.depth_range[0], ctx.depth_range[1]
        hd.axisangle[0], hd.axisangle[1] = aa0.data_ptr(), aa1.data_ptr()
        hd.translation[0], hd.translation[1] = tr0.data_ptr(), tr1.data_ptr()
        hd.invert[0], hd.invert[1] = 1, 0
        with torch.cuda.device(device):
            check(lib.cdp_photo_bwd_heads(b, h, w, num_levels, _ptr(saved), saved.numel(), _ptr(ctx.tables), _ptr(go),
                                          ctypes.byref(hd), _ptr(depth), _ptr(grad_disp), _ptr(ga0), _ptr(gt0), _ptr(ga1),
                                          _ptr(gt1), int(ctx.has_motion), _ptr(gm0), _ptr(gm1), _stream(device)),
                  "cdp_photo_bwd_heads")
        _LAUNCHES["count"] += lib.cdp_photo_bwd_launches(b, num_levels, int(ctx.has_motion))
        return (grad_disp, ga0, gt0, ga1, gt1, gm0, gm1) + (None,) * 12


def photometric_loss_from_heads(intrinsics, images: Sequence[torch.Tensor], disp: torch.Tensor,
                                pose_params: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                                noise: Optional[Sequence[torch.Tensor]], num_levels: int, alpha: float = 0.85,
                                seed: int = 0, min_depth: float = 0.1, max_depth: float = 100.0,
                                motions: Optional[Sequence[torch.Tensor]] = None,
                                noise_event: Optional[torch.cuda.Event] = None):
    """The same loss taking what the network heads emit: ``disp`` [B,1,H,W] (sigmoid disparity,
    models/depth_head.py:49-54) and ``pose_params`` = ((axis-angle, translation) for t -> t-1,
    (axis-angle, translation) for t -> t+1), each [B,3] or [B,1,3] (models/pose_head.py:47-77;
    the first pair is inverted like ``pose_head(feats, invert_pose=True)``, algos/depth.py:404-407).
    Gradients flow to ``disp`` and the four parameter tensors directly; the conversions run inside
    the pyramid / depth-gradient launches instead of as kernels (and autograd nodes) of their own.
    Returns (loss, per-level argmin maps, depth [B,1,H,W], (T0, T1) [B,4,4])."""
    target = _require_cuda_f32(images[0], "images[0]", (None, 3, None, None))
    b, _, h, w = target.shape
    source0 = _require_cuda_f32(images[1], "images[1]", (b, 3, h, w))
    source1 = _require_cuda_f32(images[2], "images[2]", (b, 3, h, w))
    disp = _require_cuda_f32(disp, "disparity_map", (b, 1, h, w))
    if len(pose_params) != 2 or any(len(pp) != 2 for pp in pose_params):
        raise ValueError("pose_params must be ((axisangle, translation), (axisangle, translation))")
    flat = []
    for k, (aa, tr) in enumerate(pose_params):
        for name, t in (("axisangle", aa), ("translation", tr)):
            if not (isinstance(t, torch.Tensor) and t.shape[0] == b and t.shape[-1] == 3 and t.numel() == 3 * b):
                raise ValueError(f"pose_params[{k}] {name} must be [B,3] or [B,1,3], got {tuple(t.shape)}")
            flat.append(_require_cuda_f32(t.reshape(b, 3), f"pose_params[{k}] {name}"))
    for name, t in [("images[1]", source0), ("images[2]", source1), ("disparity_map", disp)] + [("pose_params", t) for t in flat]:
        if t.device != target.device:
            raise RuntimeError(f"{name} is on {t.device}, images[0] on {target.device}")
    if not (min_depth > 0 and max_depth > min_depth):
        raise ValueError("need 0 < min_depth < max_depth")
    if isinstance(intrinsics, torch.Tensor):
        intrinsics = _require_cuda_f32(intrinsics.detach(), "intrinsics", (b, 4))
    else:
        intrinsics = np.ascontiguousarray(intrinsics, dtype=np.float32)
        if intrinsics.shape != (num_levels, b, 4):
            raise ValueError(f"intrinsics has shape {intrinsics.shape}, expected {(num_levels, b, 4)}")
    if not 1 <= num_levels <= _native.MAX_LEVELS:
        raise ValueError(f"num_levels must be in [1, {_native.MAX_LEVELS}], got {num_levels}")
    if noise is not None:
        if len(noise) != num_levels:
            raise ValueError(f"noise has {len(noise)} levels, expected {num_levels}")
        noise = [_require_cuda_f32(n, f"noise[{s}]", (b, 2, h >> s, w >> s)) for s, n in enumerate(noise)]
    motion0 = motion1 = None
    if motions is not None:
        if len(motions) != 2:
            raise ValueError("object_motion_maps must hold one map per source frame")
        motion0 = _require_cuda_f32(motions[0], "object_motion_maps[0]", (b, 3, h, w))
        motion1 = _require_cuda_f32(motions[1], "object_motion_maps[1]", (b, 3, h, w))
    state = PhotoState()
    loss, depth, pose0, pose1 = _PhotometricLossHeads.apply(disp, flat[0], flat[1], flat[2], flat[3], motion0, motion1,
                                                            target, source0, source1, intrinsics, noise, seed, num_levels,
                                                            alpha, min_depth, max_depth, state, noise_event)
    return loss, state.argmin, depth, (pose0, pose1)


def tiebreak_noise(batch: int, height: int, width: int, num_levels: int, seed: int, device) -> List[torch.Tensor]:
    """The draws of the built-in tie-break generator: per level [B,2,H_s,W_s] standard-normal values,
    exactly what ``photometric_loss(..., noise=None, seed=seed)`` adds (times 1e-5) to the identity
    candidates.  Passing them back as ``noise=`` reproduces that evaluation (tests, inspection)."""
    lib = _lib_for(device)
    out = []
    with torch.cuda.device(device):
        for s in range(num_levels):
            t = torch.empty((batch, 2, height >> s, width >> s), dtype=torch.float32, device=device)
            check(lib.cdp_tiebreak_noise(batch, height >> s, width >> s, s, seed & 0xFFFFFFFFFFFFFFFF, _ptr(t), _stream(device)),
                  "cdp_tiebreak_noise")
            out.append(t)
    return out


def photometric_loss(intrinsics: np.ndarray, images: Sequence[torch.Tensor], depth: torch.Tensor,
                     poses: Sequence[torch.Tensor], noise: Optional[Sequence[torch.Tensor]],
                     num_levels: int, alpha: float = 0.85, seed: int = 0,
                     motions: Optional[Sequence[torch.Tensor]] = None,
                     noise_event: Optional[torch.cuda.Event] = None
                     ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Multi-scale min-reprojection loss with identity auto-mask.

    intrinsics: float32 array [num_levels, B, 4] (fx, fy, cx, cy per level and sample), or a CUDA
    float32 tensor [B, 4] with the full-resolution values (rescaled per level inside the kernel; no
    host copy of the calibration is needed then).
    noise: per level [B,2,H_s,W_s] standard-normal tie-break draws, or None to use the kernel's
    counter-based generator with ``seed`` -- an int, or a one-element int64 CUDA tensor that the
    library reads on the device and increments after the call (fresh draws per CUDA-graph replay).  ``noise_event``: event recorded after the noise was
    written on another stream; the library makes the current stream wait on it right before the
    tile kernel, so the noise generation overlaps the pyramid kernel.
    Returns (loss, per-level argmin maps)."""
    target = _require_cuda_f32(images[0], "images[0]", (None, 3, None, None))
    b, _, h, w = target.shape
    source0 = _require_cuda_f32(images[1], "images[1]", (b, 3, h, w))
    source1 = _require_cuda_f32(images[2], "images[2]", (b, 3, h, w))
    depth = _require_cuda_f32(depth, "depth_map", (b, 1, h, w))
    pose0 = _require_cuda_f32(poses[0], "poses[0]", (b, 4, 4))
    pose1 = _require_cuda_f32(poses[1], "poses[1]", (b, 4, 4))
    if pose0.data_ptr() % 16:  # the kernels read the matrices with 16-byte loads
        pose0 = pose0.clone()
    if pose1.data_ptr() % 16:
        pose1 = pose1.clone()
    for name, t in (("images[1]", source0), ("images[2]", source1), ("depth_map", depth),
                    ("poses[0]", pose0), ("poses[1]", pose1)):
        if t.device != target.device:
            raise RuntimeError(f"{name} is on {t.device}, images[0] on {target.device}")
    if isinstance(intrinsics, torch.Tensor):
        intrinsics = _require_cuda_f32(intrinsics.detach(), "intrinsics", (b, 4))
        if intrinsics.device != target.device:
            raise RuntimeError(f"intrinsics is on {intrinsics.device}, images[0] on {target.device}")
    else:
        intrinsics = np.ascontiguousarray(intrinsics, dtype=np.float32)
        if intrinsics.shape != (num_levels, b, 4):
            raise ValueError(f"intrinsics has shape {intrinsics.shape}, expected {(num_levels, b, 4)}")
    if not 1 <= num_levels <= _native.MAX_LEVELS:
        raise ValueError(f"num_levels must be in [1, {_native.MAX_LEVELS}], got {num_levels}")
    if noise is not None:
        if len(noise) != num_levels:
            raise ValueError(f"noise has {len(noise)} levels, expected {num_levels}")
        noise = [_require_cuda_f32(n, f"noise[{s}]", (b, 2, h >> s, w >> s))
                 for s, n in enumerate(noise)]
    motion0 = motion1 = None
    if motions is not None:
        if len(motions) != 2:
            raise ValueError("object_motion_maps must hold one map per source frame")
        motion0 = _require_cuda_f32(motions[0], "object_motion_maps[0]", (b, 3, h, w))
        motion1 = _require_cuda_f32(motions[1], "object_motion_maps[1]", (b, 3, h, w))
    state = PhotoState()
    loss = _PhotometricLoss.apply(depth, pose0, pose1, motion0, motion1, target, source0, source1, intrinsics,
                                  noise, seed, num_levels, alpha, state, noise_event)
    return loss, state.argmin


class _Smoothness(torch.autograd.Function):
    """cdp_smooth_fwd / cdp_smooth_bwd."""

    @staticmethod
    def forward(ctx, disp, image):
        b, _, h, w = disp.shape
        device = disp.device
        lib = _lib_for(device)
        need_grad = ctx.needs_input_grad[0]
        saved = _bytes(lib.cdp_smooth_saved_bytes(b, h, w), device)
        loss = torch.empty(1, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(lib.cdp_smooth_fwd(_ptr(image), _ptr(disp), b, h, w, int(need_grad), _ptr(loss),
                                     _ptr(saved), saved.numel(), _stream(device)), "cdp_smooth_fwd")
        _LAUNCHES["count"] += 2
        ctx.saved_buf = saved if need_grad else None
        ctx.shape = (b, h, w)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_loss):
        b, h, w = ctx.shape
        saved = ctx.saved_buf
        device = saved.device
        lib = _lib_for(device)
        go = _require_cuda_f32(grad_loss.reshape(1), "grad_loss")
        grad_disp = torch.empty(b, 1, h, w, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(lib.cdp_smooth_bwd(_ptr(saved), saved.numel(), _ptr(go), b, h, w, _ptr(grad_disp),
                                     _stream(device)), "cdp_smooth_bwd")
        _LAUNCHES["count"] += 1
        return grad_disp, None


def smoothness_loss(target_image: torch.Tensor, disparity: torch.Tensor) -> torch.Tensor:
    image = _require_cuda_f32(target_image, "target_image", (None, 3, None, None))
    b, _, h, w = image.shape
    disp = _require_cuda_f32(disparity, "disparity_map", (b, 1, h, w))
    if h < 2 or w < 2:
        raise ValueError("edge-aware smoothness needs at least 2x2 pixels")
    return _Smoothness.apply(disp, image)


class _FlowLoss(torch.autograd.Function):
    """cdp_flow_smooth_fwd / cdp_flow_sparsity_fwd; backward = cdp_scale_fwd of the unit gradients."""

    @staticmethod
    def forward(ctx, kind, wrap_around, *maps):
        n = len(maps)
        b, c, h, w = maps[0].shape
        device = maps[0].device
        lib = _lib_for(device)
        need_grad = any(ctx.needs_input_grad[2:])
        sparsity = kind == "sparsity"
        scratch = _bytes(lib.cdp_flow_scratch_bytes(n, b * c, h, w, int(sparsity)), device)
        unit = torch.empty((n, b, c, h, w), dtype=torch.float32, device=device) if need_grad else None
        loss = torch.empty(1, dtype=torch.float32, device=device)
        ptrs = (ctypes.c_void_p * n)(*[m.data_ptr() for m in maps])
        with torch.cuda.device(device):
            if sparsity:
                check(lib.cdp_flow_sparsity_fwd(ptrs, n, b * c, h, w, _ptr(loss), _ptr(unit), _ptr(scratch),
                                                scratch.numel(), _stream(device)), "cdp_flow_sparsity_fwd")
            else:
                check(lib.cdp_flow_smooth_fwd(ptrs, n, b * c, h, w, int(wrap_around), _ptr(loss), _ptr(unit),
                                              _ptr(scratch), scratch.numel(), _stream(device)), "cdp_flow_smooth_fwd")
        _LAUNCHES["count"] += 3 if sparsity else 2
        ctx.unit = unit
        return loss[0]

    @staticmethod
    def backward(ctx, grad_loss):
        unit = ctx.unit
        device = unit.device
        lib = _lib_for(device)
        go = _require_cuda_f32(grad_loss.reshape(1), "grad_loss")
        out = torch.empty_like(unit)
        with torch.cuda.device(device):
            check(lib.cdp_scale_fwd(_ptr(unit), _ptr(go), unit.numel(), _ptr(out), _stream(device)), "cdp_scale_fwd")
        _LAUNCHES["count"] += 1
        return (None, None) + tuple(out[i] if ctx.needs_input_grad[2 + i] else None for i in range(out.shape[0]))


def _check_flow_maps(flow_maps) -> List[torch.Tensor]:
    if isinstance(flow_maps, torch.Tensor):
        raise TypeError("flow_maps must be a tuple of tensors (one per source frame), as in the reference")
    maps = list(flow_maps)
    if not 1 <= len(maps) <= _native.MAX_FLOW_MAPS:
        raise ValueError(f"between 1 and {_native.MAX_FLOW_MAPS} flow maps are supported, got {len(maps)}")
    first = _require_cuda_f32(maps[0], "flow_maps[0]", (None, None, None, None))
    out = [first] + [_require_cuda_f32(m, f"flow_maps[{i}]", tuple(first.shape)) for i, m in enumerate(maps[1:], 1)]
    for i, m in enumerate(out):
        if m.device != first.device:
            raise RuntimeError(f"flow_maps[{i}] is on {m.device}, flow_maps[0] on {first.device}")
    return out


def flow_smoothness_loss(flow_maps, wrap_around: bool = True) -> torch.Tensor:
    """FlowSmoothnessLoss.__call__ (/root/reference/algos/depth.py:29-34)."""
    maps = _check_flow_maps(flow_maps)
    if not wrap_around and (maps[0].shape[2] < 2 or maps[0].shape[3] < 2):
        raise ValueError("without wrap-around the maps must be at least 2x2")
    return _FlowLoss.apply("smooth", bool(wrap_around), *maps)


def flow_sparsity_loss(flow_maps) -> torch.Tensor:
    """FlowSparsityLoss.__call__ (/root/reference/algos/depth.py:46-51)."""
    return _FlowLoss.apply("sparsity", True, *_check_flow_maps(flow_maps))


def depth_metrics(depth_gt: torch.Tensor, depth_pred: torch.Tensor, min_depth: float, max_depth: float,
                  use_gt_scale: bool, garg_crop: bool = False, labels: Optional[torch.Tensor] = None,
                  class_id: int = 0) -> torch.Tensor:
    """cdp_depth_metrics_fwd.  depth_gt / depth_pred: [B,1,H,W] fp32 CUDA.  Without ``labels`` every
    image is a unit and the result is the batch mean; with ``labels`` ([B,1,H,W] int64) the whole
    batch restricted to ``class_id`` is one unit.  Returns a device tensor of 8 floats: d_a1, d_a2,
    d_a3, d_rmse, d_rmse_log, d_abs_rel, d_sq_rel, number of units with ground truth."""
    gt = _require_cuda_f32(depth_gt, "depth_gt", (None, 1, None, None))
    b, _, h, w = gt.shape
    pred = _require_cuda_f32(depth_pred.detach(), "depth_pred", (b, 1, h, w))
    device = gt.device
    if labels is not None:
        if not labels.is_cuda or labels.dtype != torch.int64 or tuple(labels.shape) != (b, 1, h, w):
            raise ValueError("labels must be an int64 CUDA tensor of shape [B,1,H,W]")
        labels = labels.contiguous()
        units, count = 1, b * h * w
        if garg_crop:
            raise NotImplementedError("the per-class metrics of the reference do not apply the Garg crop")
    else:
        units, count = b, h * w
    lib = _lib_for(device)
    scratch = _bytes(lib.cdp_depth_metrics_scratch_bytes(units, count), device)
    out = torch.empty(8, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib.cdp_depth_metrics_fwd(_ptr(gt), _ptr(pred), _ptr(labels), int(class_id), units, count, h, w,
                                        int(garg_crop), float(min_depth), float(max_depth), int(use_gt_scale),
                                        _ptr(out), _ptr(scratch), scratch.numel(), _stream(device)),
              "cdp_depth_metrics_fwd")
    _LAUNCHES["count"] += 6
    return out


def warp_c2c(src: torch.Tensor, k_src: np.ndarray, k_tgt: np.ndarray, out_hw: Tuple[int, int],
             depth_val: float = 1.0, interp_mode: str = "bilinear", padding_mode: str = "border") -> torch.Tensor:
    """cdp_warp_c2c_fwd: [B,C,Hs,Ws] (fp32 or fp64, CUDA) -> fp64 [B,C,Ht,Wt]; k_* are [B,4] host
    arrays (fx, fy, cx, cy) of the source / target cameras.  No gradient (as in the reference,
    where the warp is a data augmentation on detached tensors)."""
    if interp_mode not in ("bilinear", "nearest"):
        raise NotImplementedError(f"interp_mode {interp_mode!r}: only 'bilinear' and 'nearest' are implemented")
    if padding_mode not in ("border", "zeros"):
        raise NotImplementedError(f"padding_mode {padding_mode!r}: only 'border' and 'zeros' are implemented")
    if not isinstance(src, torch.Tensor) or not src.is_cuda:
        raise RuntimeError("in_src must be a CUDA tensor; codeps_b200 runs on CUDA only (no CPU fallback)")
    if src.dim() != 4:
        raise ValueError(f"in_src must be [B,C,H,W], got {tuple(src.shape)}")
    if src.dtype not in (torch.float32, torch.float64):
        src = src.double()  # labels / masks: the reference converts with .double() too (mixup.py:226)
    src = src.detach().contiguous()
    b, c, hs, ws = src.shape
    ks = np.ascontiguousarray(k_src, dtype=np.float64).reshape(-1, 4)
    kt = np.ascontiguousarray(k_tgt, dtype=np.float64).reshape(-1, 4)
    if ks.shape[0] != b or kt.shape[0] != b:
        raise ValueError(f"need one source and one target camera per sample: {ks.shape[0]} / {kt.shape[0]} for batch {b}")
    ht, wt = int(out_hw[0]), int(out_hw[1])
    device = src.device
    lib = _lib_for(device)
    out = torch.empty((b, c, ht, wt), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib.cdp_warp_c2c_fwd(_ptr(src), int(src.dtype == torch.float64), b, c, hs, ws, ht, wt,
                                   ks.ctypes.data_as(ctypes.c_void_p), kt.ctypes.data_as(ctypes.c_void_p),
                                   float(depth_val), int(interp_mode == "nearest"), int(padding_mode == "zeros"),
                                   _ptr(out), _stream(device)), "cdp_warp_c2c_fwd")
    _LAUNCHES["count"] += (b + _native.MAX_BATCH_PER_LAUNCH - 1) // _native.MAX_BATCH_PER_LAUNCH
    return out


class _WarpImage(torch.autograd.Function):
    """cdp_warp_image_fwd / cdp_warp_image_bwd."""

    @staticmethod
    def forward(ctx, src, depth, pose, motion, intrinsics, mode):
        if ctx.needs_input_grad[0]:
            raise RuntimeError("ImageWarper: gradients w.r.t. the source image are not supported "
                               "(the loss never needs them); detach the source image")
        b, c, h, w = src.shape
        device = src.device
        lib = _lib_for(device)
        out = torch.empty_like(src)
        with torch.cuda.device(device):
            check(lib.cdp_warp_image_fwd(_ptr(src), c, _ptr(depth), _ptr(pose), _ptr(motion),
                                         ctypes.c_void_p(intrinsics.ctypes.data), b, h, w, mode, _ptr(out),
                                         _stream(device)), "cdp_warp_image_fwd")
        _LAUNCHES["count"] += (b + _native.MAX_BATCH_PER_LAUNCH - 1) // _native.MAX_BATCH_PER_LAUNCH
        ctx.save_for_backward(src, depth, pose, motion if motion is not None else torch.empty(0))
        ctx.intrinsics = intrinsics
        ctx.mode = mode
        ctx.has_motion = motion is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        src, depth, pose, motion = ctx.saved_tensors
        motion = motion if ctx.has_motion else None
        if ctx.mode != 0:
            # F.grid_sample(mode="nearest") has a zero coordinate gradient (ATen returns zeros)
            return (None, torch.zeros_like(depth), torch.zeros_like(pose),
                    torch.zeros_like(motion) if motion is not None else None, None, None)
        b, c, h, w = src.shape
        device = src.device
        lib = _lib_for(device)
        grad_out = _require_cuda_f32(grad_out, "grad_output", (b, c, h, w))
        grad_depth = torch.empty_like(depth)
        grad_pose = torch.empty_like(pose)
        grad_motion = torch.empty_like(motion) if motion is not None else None
        scratch = _bytes(lib.cdp_warp_bwd_scratch_bytes(b, h, w), device)
        with torch.cuda.device(device):
            check(lib.cdp_warp_image_bwd(_ptr(grad_out), _ptr(src), c, _ptr(depth), _ptr(pose), _ptr(motion),
                                         ctypes.c_void_p(ctx.intrinsics.ctypes.data), b, h, w,
                                         _ptr(grad_depth), _ptr(grad_pose), _ptr(grad_motion), _ptr(scratch),
                                         scratch.numel(), _stream(device)), "cdp_warp_image_bwd")
        _LAUNCHES["count"] += 1 + (b + _native.MAX_BATCH_PER_LAUNCH - 1) // _native.MAX_BATCH_PER_LAUNCH
        return None, grad_depth, grad_pose, grad_motion, None, None


def _check_warp_inputs(depth, pose, motion, intrinsics):
    depth = _require_cuda_f32(depth, "batch_depth_map", (None, 1, None, None))
    b, _, h, w = depth.shape
    pose = _require_cuda_f32(pose, "T", (b, 4, 4))
    if motion is not None:
        motion = _require_cuda_f32(motion, "object_motion_map", (b, 3, h, w))
    intrinsics = np.ascontiguousarray(intrinsics, dtype=np.float32)
    if intrinsics.shape != (b, 4):
        raise ValueError(f"intrinsics has shape {intrinsics.shape}, expected {(b, 4)}")
    if h < 2 or w < 2:
        raise ValueError("warping needs at least 2x2 pixels")
    return depth, pose, motion, intrinsics


def warp_image(src, depth, pose, intrinsics, mode: str = "bilinear", motion=None) -> torch.Tensor:
    if mode not in ("bilinear", "nearest"):
        raise ValueError(f"interp_mode must be 'bilinear' or 'nearest', got {mode!r}")
    depth, pose, motion, intrinsics = _check_warp_inputs(depth, pose, motion, intrinsics)
    b, _, h, w = depth.shape
    src = _require_cuda_f32(src, "batch_src_img", (b, None, h, w))
    return _WarpImage.apply(src, depth, pose, motion, intrinsics, 0 if mode == "bilinear" else 1)


def warp_grid(depth, pose, intrinsics, motion=None) -> torch.Tensor:
    """Normalised sampling grid [B,H,W,2] (forward only)."""
    depth, pose, motion, intrinsics = _check_warp_inputs(depth, pose, motion, intrinsics)
    if any(t is not None and t.requires_grad for t in (depth, pose, motion)) and torch.is_grad_enabled():
        raise RuntimeError("CoordinateWarper: the stand-alone grid output is forward-only; use "
                           "ImageWarper / ReconstructionLoss for differentiable warping")
    b, _, h, w = depth.shape
    device = depth.device
    lib = _lib_for(device)
    grid = torch.empty(b, h, w, 2, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib.cdp_warp_grid_fwd(_ptr(depth), _ptr(pose), _ptr(motion),
                                    ctypes.c_void_p(intrinsics.ctypes.data), b, h, w, _ptr(grid),
                                    _stream(device)), "cdp_warp_grid_fwd")
    _LAUNCHES["count"] += (b + _native.MAX_BATCH_PER_LAUNCH - 1) // _native.MAX_BATCH_PER_LAUNCH
    return grid


class _Ssim(torch.autograd.Function):
    """cdp_ssim_fwd / cdp_ssim_bwd."""

    @staticmethod
    def forward(ctx, x, y):
        b, c, h, w = x.shape
        device = x.device
        lib = _lib_for(device)
        out = torch.empty_like(x)
        with torch.cuda.device(device):
            check(lib.cdp_ssim_fwd(_ptr(x), _ptr(y), b * c, h, w, _ptr(out), _stream(device)), "cdp_ssim_fwd")
        _LAUNCHES["count"] += 1
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, y = ctx.saved_tensors
        b, c, h, w = x.shape
        device = x.device
        lib = _lib_for(device)
        grad_out = _require_cuda_f32(grad_out, "grad_output", (b, c, h, w))
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        scratch = _bytes(lib.cdp_ssim_bwd_scratch_bytes(b * c, h, w), device)
        with torch.cuda.device(device):
            check(lib.cdp_ssim_bwd(_ptr(grad_out), _ptr(x), _ptr(y), b * c, h, w, _ptr(gx), _ptr(gy),
                                   _ptr(scratch), scratch.numel(), _stream(device)), "cdp_ssim_bwd")
        _LAUNCHES["count"] += 2
        return gx, gy


def ssim_loss_map(src_img: torch.Tensor, target_img: torch.Tensor) -> torch.Tensor:
    x = _require_cuda_f32(src_img, "src_img", (None, None, None, None))
    y = _require_cuda_f32(target_img, "target_img", tuple(x.shape))
    if x.shape[2] < 2 or x.shape[3] < 2:
        raise ValueError("SSIM with reflection padding needs at least 2x2 pixels")
    return _Ssim.apply(x, y)


class _PoseMatrix(torch.autograd.Function):
    """cdp_pose_fwd / cdp_pose_bwd."""

    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        b = axisangle.shape[0]
        device = axisangle.device
        lib = _lib_for(device)
        out = torch.empty(b, 4, 4, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(lib.cdp_pose_fwd(_ptr(axisangle), _ptr(translation), b, int(invert), _ptr(out), _stream(device)),
                  "cdp_pose_fwd")
        _LAUNCHES["count"] += 1
        ctx.save_for_backward(axisangle, translation)
        ctx.invert = bool(invert)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        axisangle, translation = ctx.saved_tensors
        b = axisangle.shape[0]
        device = axisangle.device
        lib = _lib_for(device)
        grad_out = _require_cuda_f32(grad_out, "grad_output", (b, 4, 4))
        ga, gt = torch.empty_like(axisangle), torch.empty_like(translation)
        with torch.cuda.device(device):
            check(lib.cdp_pose_bwd(_ptr(grad_out), _ptr(axisangle), _ptr(translation), b, int(ctx.invert), _ptr(ga),
                                   _ptr(gt), _stream(device)), "cdp_pose_bwd")
        _LAUNCHES["count"] += 1
        return ga, gt, None


def pose_matrix(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False) -> torch.Tensor:
    """6-DoF -> [B,4,4].  axisangle / translation: [B,3] or [B,1,3] (the pose head's layout)."""
    shape_ok = lambda t: t.dim() in (2, 3) and t.shape[-1] == 3 and (t.dim() == 2 or t.shape[1] == 1)
    if not (isinstance(axisangle, torch.Tensor) and isinstance(translation, torch.Tensor)):
        raise TypeError("axisangle and translation must be tensors")
    if not shape_ok(axisangle) or not shape_ok(translation) or axisangle.shape[0] != translation.shape[0]:
        raise ValueError(f"expected [B,3] or [B,1,3] inputs, got {tuple(axisangle.shape)} and {tuple(translation.shape)}")
    b = axisangle.shape[0]
    aa = _require_cuda_f32(axisangle.reshape(b, 3), "axisangle")
    tr = _require_cuda_f32(translation.reshape(b, 3), "translation")
    return _PoseMatrix.apply(aa, tr, bool(invert))


class _DispToDepth(torch.autograd.Function):
    """cdp_disp_to_depth_fwd / cdp_disp_to_depth_bwd."""

    @staticmethod
    def forward(ctx, disp, min_depth, max_depth):
        device = disp.device
        lib = _lib_for(device)
        depth = torch.empty_like(disp)
        with torch.cuda.device(device):
            check(lib.cdp_disp_to_depth_fwd(_ptr(disp), disp.numel(), float(min_depth), float(max_depth), _ptr(depth),
                                            _stream(device)), "cdp_disp_to_depth_fwd")
        _LAUNCHES["count"] += 1
        ctx.save_for_backward(depth)
        ctx.range = (float(min_depth), float(max_depth))
        return depth

    @staticmethod
    def backward(ctx, grad_depth):
        (depth,) = ctx.saved_tensors
        device = depth.device
        lib = _lib_for(device)
        grad_depth = _require_cuda_f32(grad_depth, "grad_output", tuple(depth.shape))
        grad_disp = torch.empty_like(depth)
        with torch.cuda.device(device):
            check(lib.cdp_disp_to_depth_bwd(_ptr(grad_depth), _ptr(depth), depth.numel(), ctx.range[0], ctx.range[1],
                                            _ptr(grad_disp), _stream(device)), "cdp_disp_to_depth_bwd")
        _LAUNCHES["count"] += 1
        return grad_disp, None, None


def disp_to_depth(disp: torch.Tensor, min_depth: float = 0.1, max_depth: float = 100.0) -> torch.Tensor:
    disp = _require_cuda_f32(disp, "disp")
    if disp.numel() == 0:
        raise ValueError("empty disparity tensor")
    if not (min_depth > 0 and max_depth > min_depth):
        raise ValueError("need 0 < min_depth < max_depth")
    return _DispToDepth.apply(disp, min_depth, max_depth)
