"""Drop-in ``ImageWarper`` / ``CoordinateWarper`` backed by the CUDA kernels.

Same constructor and ``forward`` signatures as /root/reference/misc/image_warper.py:90-184
(``ImageWarper(img_width, img_height, device)``, ``.coordinate_warper``), so code that builds a
warper per pyramid level (/root/reference/algos/depth.py:215) or calls it directly
(/root/reference/algos/semantic_seg.py:70-144) keeps working.  Back-projection, the SE(3)
transform, projection and the bilinear / nearest gather run in one kernel
(cdp_warp_image_fwd); backward goes to depth, pose and the optional object-motion map.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
from torch import nn

from . import ops
from .camera import CameraModel


def intrinsics_table(camera_models: Sequence[CameraModel]) -> np.ndarray:
    """[B,4] fp32 (fx, fy, cx, cy), rounded the way the reference rounds them on use."""
    return np.asarray([[np.float32(cam.intrinsics[k]) for k in ("fx", "fy", "cx", "cy")]
                       for cam in camera_models], dtype=np.float32).reshape(len(camera_models), 4)


class CoordinateWarper(nn.Module):
    """depth + T (+ object motion) -> normalised sampling grid [B,H,W,2]
    (/root/reference/misc/image_warper.py:90-144)."""

    def __init__(self, img_width: int, img_height: int, device: torch.device):
        super().__init__()
        self.img_width = img_width
        self.img_height = img_height
        self.device = device

    def _check_size(self, depth: torch.Tensor):
        assert depth.dim() == 4, f"The input batch of depth maps has {depth.dim()} dimensions which is != 4"
        assert depth.size(1) == 1, f"The input batch of depth maps has {depth.size(1)} channels which is != 1"
        if depth.shape[2] != self.img_height or depth.shape[3] != self.img_width:
            raise ValueError(f"depth map is {depth.shape[3]}x{depth.shape[2]}, warper was built for "
                             f"{self.img_width}x{self.img_height}")

    def forward(self, batch_camera_models: List[CameraModel], batch_depth_map, T, object_motion_map=None):
        self._check_size(batch_depth_map)
        return ops.warp_grid(batch_depth_map, T, intrinsics_table(batch_camera_models), object_motion_map)


class ImageWarper(nn.Module):
    """Warp ``batch_src_img`` into the target view (/root/reference/misc/image_warper.py:147-184):
    grid_sample(bilinear | nearest, padding_mode="border", align_corners=True) at the reprojected
    coordinates."""

    def __init__(self, img_width: int, img_height: int, device: torch.device):
        super().__init__()
        self.coordinate_warper = CoordinateWarper(img_width, img_height, device)

    def forward(self, batch_camera_models: List[CameraModel], batch_src_img, batch_depth_map, T,
                interp_mode="bilinear", object_motion_map=None):
        assert batch_src_img.dim() == 4, \
            f"The input batch of source images has {batch_src_img.dim()} dimensions which is != 4"
        self.coordinate_warper._check_size(batch_depth_map)
        return ops.warp_image(batch_src_img, batch_depth_map, T, intrinsics_table(batch_camera_models),
                              mode=interp_mode, motion=object_motion_map)
