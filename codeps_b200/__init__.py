"""codeps_b200 -- B200-native photometric reprojection loss for CoDEPS.

Drop-in replacements for the reference's hot-path classes (``CameraModel``, ``ImageWarper``,
``SSIMLoss``, ``ReconstructionLoss``, ``EdgeAwareSmoothnessLoss``) on top of hand-written
sm_100a CUDA kernels behind a C ABI (include/codeps_photo.h).  See DESIGN.md.
"""
from .camera import CameraModel
from .evaluator import DepthEvaluator
from .heads import disp_to_depth, transformation_from_parameters
from .install import install, uninstall
from .losses import (EdgeAwareSmoothnessLoss, FlowSmoothnessLoss, FlowSparsityLoss, ReconstructionLoss,
                     SSIMLoss)
from .mixup import warp_c2c
from .warper import CoordinateWarper, ImageWarper

__all__ = ["CameraModel", "ImageWarper", "CoordinateWarper", "SSIMLoss", "ReconstructionLoss",
           "EdgeAwareSmoothnessLoss", "FlowSmoothnessLoss", "FlowSparsityLoss", "install", "uninstall", "transformation_from_parameters", "disp_to_depth", "warp_c2c", "DepthEvaluator"]
__version__ = "0.1.0"
