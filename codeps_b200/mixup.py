"""Camera-to-camera warp of the mix-up augmentation, backed by the CUDA kernel.

``warp_c2c`` has the signature of ``Mixup.warp_c2c``
(/root/reference/datasets/mixup.py:211-229): the source image (or label / instance map) is
re-rendered through the target camera's intrinsics at a constant depth -- viewing rays of the
target camera, projection with the source camera, ``F.grid_sample(align_corners=True)`` in fp64.
One kernel instead of ~25 element-wise launches and a python loop over the batch.

``codeps_b200.install()`` binds it over ``datasets.mixup.Mixup.warp_c2c`` (used by
``Mixup.embed_wsrc2tgt`` and the geometric augmentation at mixup.py:430-441).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ops
from .camera import CameraModel


def _intrinsics64(camera_models: List[CameraModel]) -> np.ndarray:
    # python floats stay fp64, np.float32 values (CameraModel.from_tensor) convert exactly
    return np.asarray([[float(cam.intrinsics[k]) for k in ("fx", "fy", "cx", "cy")] for cam in camera_models],
                      dtype=np.float64).reshape(len(camera_models), 4)


def warp_c2c(cam_model_src: List[CameraModel], cam_model_tgt: List[CameraModel], in_src: torch.Tensor,
             in_tgt: torch.Tensor, depth_val: Optional[float] = 1, interp_mode="bilinear",
             padding_mode="border") -> torch.Tensor:
    if in_src.dim() == 3:  # label maps come as [B,H,W] (mixup.py:215-216)
        in_src = in_src.unsqueeze(1)
    if len(cam_model_src) != in_src.shape[0] or len(cam_model_tgt) != in_src.shape[0]:
        raise ValueError("one source and one target camera model per sample are required")
    if in_tgt.shape[0] != in_src.shape[0]:
        raise ValueError(f"batch sizes differ: in_src {in_src.shape[0]}, in_tgt {in_tgt.shape[0]}")
    return ops.warp_c2c(in_src, _intrinsics64(cam_model_src), _intrinsics64(cam_model_tgt),
                        (in_tgt.shape[-2], in_tgt.shape[-1]), 1.0 if depth_val is None else depth_val,
                        interp_mode, padding_mode)
