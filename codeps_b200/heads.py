"""CUDA versions of the two conversions that sit between the network heads and the loss.

``transformation_from_parameters`` mirrors ``PoseHead.transformation_from_parameters``
(/root/reference/models/pose_head.py:56-77, with ``rot_from_axisangle`` :80-119 and
``get_translation_matrix`` :122-137): same argument order and shapes ([B,1,3] axis-angle and
translation as the pose head slices them, pose_head.py:47-52), one kernel forward and one backward
instead of ~40 tiny element kernels.  ``disp_to_depth`` mirrors ``DepthHead.disp_to_depth``
(/root/reference/models/depth_head.py:49-54).  SURVEY.md section 8f, row 1.
"""
from __future__ import annotations

from torch import Tensor

from . import ops


def transformation_from_parameters(axisangle: Tensor, translation: Tensor, invert: bool = False) -> Tensor:
    """(axis-angle, translation) -> [B,4,4]; ``invert`` gives R^T T(-t), the t -> t-1 pose."""
    return ops.pose_matrix(axisangle, translation, invert)


def disp_to_depth(disp: Tensor, min_depth: float = 0.1, max_depth: float = 100) -> Tensor:
    """Sigmoid disparity -> depth in [min_depth, max_depth]."""
    return ops.disp_to_depth(disp, min_depth, max_depth)
