"""Seeded synthetic frame triplets shaped like the datasets CoDEPS trains on.

There are no datasets in the build or bench environment, so tests and ``bench.py`` feed the
loss with textures that behave like real input to it: ImageNet-normalised RGB
(/root/reference/datasets/preprocessing.py:12-18), a sigmoid-range disparity converted to
depth the way ``DepthHead.disp_to_depth`` does (/root/reference/models/depth_head.py:49-54) and
4x4 poses assembled like ``PoseHead.transformation_from_parameters``
(/root/reference/models/pose_head.py:56-137).  The neighbouring frames are the target frame
shifted by a few pixels, and the pose translation explains most of that shift, so the
min-reprojection picks a healthy mix of reprojection and identity (auto-mask) pixels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch
import torch.nn.functional as F

from .camera import CameraModel

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

# (width, height, fx, fy, cx, cy) -- typical calibration values for the three datasets
# (SURVEY.md section 8d); they are not stored in the reference.
PRESETS = {
    "cityscapes": (1024, 512, 1131.26, 1132.65, 548.49, 256.5685),
    "kitti360": (1408, 376, 552.554261, 552.554261, 682.049453, 238.769549),
    "kitti360_cfg": (1408, 384, 552.554261, 564.31, 682.049453, 243.85),
    "semkitti": (1280, 384, 738.23, 733.85, 628.40, 190.04),
}


def disp_to_depth(disp: torch.Tensor, min_depth: float = 0.1, max_depth: float = 100.0):
    """Sigmoid disparity in (0,1) -> metric depth in [min_depth, max_depth]."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    return 1.0 / (lo + (hi - lo) * disp)


def _rodrigues(axisangle: torch.Tensor) -> torch.Tensor:
    """[B,3] axis-angle -> [B,4,4] homogeneous rotation."""
    theta = axisangle.norm(dim=1, keepdim=True)
    axis = axisangle / (theta + 1e-7)
    c, s = torch.cos(theta)[:, 0], torch.sin(theta)[:, 0]
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    v = 1 - c
    rot = torch.zeros(axisangle.shape[0], 4, 4, dtype=axisangle.dtype, device=axisangle.device)
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 0, 2] = x * x * v + c, x * y * v - z * s, z * x * v + y * s
    rot[:, 1, 0], rot[:, 1, 1], rot[:, 1, 2] = x * y * v + z * s, y * y * v + c, y * z * v - x * s
    rot[:, 2, 0], rot[:, 2, 1], rot[:, 2, 2] = z * x * v - y * s, y * z * v + x * s, z * z * v + c
    rot[:, 3, 3] = 1
    return rot


def pose_matrix(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False):
    """6-DoF (axis-angle [B,3], translation [B,3]) -> [B,4,4]; ``invert`` gives the inverse
    motion (R^T applied after translating by -t), as used for the t -> t-1 pose."""
    rot = _rodrigues(axisangle)
    t = -translation if invert else translation
    trans = torch.eye(4, dtype=t.dtype, device=t.device).repeat(t.shape[0], 1, 1)
    trans[:, :3, 3] = t
    if invert:
        return rot.transpose(1, 2) @ trans
    return trans @ rot


@dataclass
class TripletBatch:
    """One batch of the hot path's inputs (all fp32 NCHW, contiguous)."""
    images: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]  # t, t-1, t+1   [B,3,H,W]
    disp: torch.Tensor  # [B,1,H,W] in (0,1)
    depth: torch.Tensor  # [B,1,H,W]
    poses: Tuple[torch.Tensor, torch.Tensor]  # t->t-1, t->t+1   [B,4,4]
    intrinsics: torch.Tensor  # [B,4] fx,fy,cx,cy
    width: int
    height: int

    def camera_models(self) -> List[CameraModel]:
        return [CameraModel.from_tensor(self.width, self.height, k) for k in self.intrinsics]

    def to(self, device, non_blocking: bool = False) -> "TripletBatch":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return TripletBatch(tuple(mv(i) for i in self.images), mv(self.disp), mv(self.depth),
                            tuple(mv(p) for p in self.poses), self.intrinsics, self.width,
                            self.height)

    def pin(self) -> "TripletBatch":
        pn = lambda t: t.pin_memory()
        return TripletBatch(tuple(pn(i) for i in self.images), pn(self.disp), pn(self.depth),
                            tuple(pn(p) for p in self.poses), self.intrinsics, self.width,
                            self.height)

    def nbytes(self) -> int:
        ts = list(self.images) + [self.disp, self.depth] + list(self.poses)
        return sum(t.numel() * t.element_size() for t in ts)


def _smooth_field(shape, gen, cells: int = 8):
    b, c, h, w = shape
    coarse = torch.rand(b, c, h // cells + 2, w // cells + 2, generator=gen)
    return F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)


def make_batch(batch: int, width: int, height: int, intrinsics, seed: int = 0, shift_px: int = 3,
               flip_every_other: bool = False, depth_range: str = "near",
               static_frac: float = 0.3) -> TripletBatch:
    """Build one seeded batch on the CPU.

    ``intrinsics`` is (fx, fy, cx, cy) for this image size.  ``flip_every_other`` mirrors the
    principal point of odd samples the way the flip augmentation does
    (/root/reference/datasets/preprocessing.py:47-52), giving per-sample intrinsics.
    ``static_frac``: the leftmost fraction of the neighbouring frames shows the target unshifted
    (content moving with the camera), which is what the identity auto-mask is there to catch.
    ``depth_range``: "near" keeps depth around 1 so that the translation explains the pixel
    shift (reprojection-dominated); "wide" spans most of the sigmoid range (identity-dominated).
    """
    gen = torch.Generator().manual_seed(seed)
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)

    tgt = _smooth_field((batch, 3, height, width), gen) + 0.05 * torch.rand(
        batch, 3, height, width, generator=gen)
    prev = torch.roll(tgt, shift_px, dims=3) + 0.02 * torch.randn(batch, 3, height, width,
                                                                   generator=gen)
    nxt = torch.roll(tgt, -shift_px, dims=3) + 0.02 * torch.randn(batch, 3, height, width,
                                                                   generator=gen)
    static_cols = int(width * static_frac)
    if static_cols > 0:
        prev[..., :static_cols] = tgt[..., :static_cols] + 0.02 * torch.randn(
            batch, 3, height, static_cols, generator=gen)
        nxt[..., :static_cols] = tgt[..., :static_cols] + 0.02 * torch.randn(
            batch, 3, height, static_cols, generator=gen)
    images = tuple(((im - mean) / std).contiguous() for im in (tgt, prev, nxt))

    field = _smooth_field((batch, 1, height, width), gen, cells=16)
    if depth_range == "near":
        disp = 0.08 + 0.04 * field
    else:
        disp = 0.02 + 0.9 * field
    disp = disp.contiguous()
    depth = disp_to_depth(disp).contiguous()

    aa = 1e-3 * torch.randn(batch, 2, 3, generator=gen)
    jitter = 2e-4 * torch.randn(batch, 2, 3, generator=gen)
    tx = shift_px / fx
    # parameters as a pose head would emit them; the t-1 pose is built inverted
    t_prev = jitter[:, 0] + torch.tensor([-tx, 0.0, 0.0])
    t_next = jitter[:, 1] + torch.tensor([-tx, 0.0, 0.0])
    poses = (pose_matrix(aa[:, 0], t_prev, invert=True).contiguous(),
             pose_matrix(aa[:, 1], t_next, invert=False).contiguous())

    k = torch.tensor([fx, fy, cx, cy], dtype=torch.float32).repeat(batch, 1)
    if flip_every_other:
        k[1::2, 2] = width - k[1::2, 2] - 1
    return TripletBatch(images, disp, depth, poses, k, width, height)


def make_preset_batch(name: str, batch: int, seed: int = 0, **kw) -> TripletBatch:
    w, h, fx, fy, cx, cy = PRESETS[name]
    return make_batch(batch, w, h, (fx, fy, cx, cy), seed=seed, **kw)


def level_sizes(width: int, height: int, num_scales: int):
    """(W_s, H_s) per level: integer halving as in /root/reference/algos/depth.py:211-214."""
    return [(width // 2**s, height // 2**s) for s in range(num_scales)]


def algorithmic_bytes(width: int, height: int, num_scales: int = 5) -> int:
    """A_alg of SURVEY.md section 8d / BASELINE.md section 3: bytes per frame triplet, fwd+bwd."""
    n = width * height
    s0 = sum(w * h for w, h in level_sizes(width, height, num_scales))
    s1 = s0 - n
    return (41 * s0 + 40 * s1 + 8 * n) + (45 * s0 + 4 * s1 + 8 * n)


__all__ = [
    "PRESETS", "TripletBatch", "make_batch", "make_preset_batch", "pose_matrix", "disp_to_depth",
    "level_sizes", "algorithmic_bytes", "IMAGENET_MEAN", "IMAGENET_STD"
]
_ = math
