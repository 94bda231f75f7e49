"""Swap the CUDA-backed classes into an importable CoDEPS checkout.

    import codeps_b200
    codeps_b200.install()          # before codeps.model_setup.gen_models(...) runs

The reference looks its loss classes up as module attributes when ``gen_models`` executes
(/root/reference/codeps/model_setup.py:3-17,63-85) and ``algos/depth.py`` imports
``CameraModel`` / ``ImageWarper`` from ``misc`` (depth.py:9), so rebinding those names is all a
drop-in needs; ``scripts/train_codeps.py`` and ``scripts/adapt_codeps.py`` stay untouched.
"""
from __future__ import annotations

import importlib
import sys

from .camera import CameraModel
from .losses import (EdgeAwareSmoothnessLoss, FlowSmoothnessLoss, FlowSparsityLoss, ReconstructionLoss,
                     SSIMLoss)
from .evaluator import DepthEvaluator
from .heads import disp_to_depth, transformation_from_parameters
from .mixup import warp_c2c
from .warper import CoordinateWarper, ImageWarper

_PATCHES = {
    "misc.camera_model": {"CameraModel": CameraModel},
    "misc.image_warper": {"ImageWarper": ImageWarper, "CoordinateWarper": CoordinateWarper},
    "misc": {"CameraModel": CameraModel, "ImageWarper": ImageWarper},
    "algos.depth": {"CameraModel": CameraModel, "ImageWarper": ImageWarper, "SSIMLoss": SSIMLoss,
                    "ReconstructionLoss": ReconstructionLoss,
                    "EdgeAwareSmoothnessLoss": EdgeAwareSmoothnessLoss,
                    "FlowSmoothnessLoss": FlowSmoothnessLoss, "FlowSparsityLoss": FlowSparsityLoss,
                    "DepthEvaluator": DepthEvaluator},
    "algos": {"SSIMLoss": SSIMLoss, "ReconstructionLoss": ReconstructionLoss,
              "EdgeAwareSmoothnessLoss": EdgeAwareSmoothnessLoss,
              "FlowSmoothnessLoss": FlowSmoothnessLoss, "FlowSparsityLoss": FlowSparsityLoss},
    "codeps.model_setup": {"SSIMLoss": SSIMLoss, "ReconstructionLoss": ReconstructionLoss,
                           "EdgeAwareSmoothnessLoss": EdgeAwareSmoothnessLoss,
                           "FlowSmoothnessLoss": FlowSmoothnessLoss, "FlowSparsityLoss": FlowSparsityLoss,
                           "DepthEvaluator": DepthEvaluator},
    "codeps.online_adap": {"CameraModel": CameraModel},
    "eval.depth": {"DepthEvaluator": DepthEvaluator},
    "eval": {"DepthEvaluator": DepthEvaluator},
}
# static methods rebound on a class: (module, class) -> {name: function}
_METHOD_PATCHES = {
    ("datasets.mixup", "Mixup"): {"warp_c2c": warp_c2c},
    # the two conversions between the network heads and the loss (models/pose_head.py:56-77,
    # models/depth_head.py:49-54): one kernel forward / backward each instead of ~40 element kernels
    ("models.pose_head", "PoseHead"): {"transformation_from_parameters": transformation_from_parameters},
    ("models.depth_head", "DepthHead"): {"disp_to_depth": disp_to_depth},
}
_originals = {}


def install(import_missing: bool = True) -> list:
    """Rebind the hot-path classes in every CoDEPS module that is (or can be) imported.
    Returns the list of ``module.attribute`` names that were patched."""
    patched = []
    for mod_name, names in _PATCHES.items():
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:  # CoDEPS not on sys.path, or an optional dependency is missing
                continue
        if mod is None:
            continue
        for attr, obj in names.items():
            if hasattr(mod, attr):
                _originals.setdefault((mod_name, attr), getattr(mod, attr))
                setattr(mod, attr, obj)
                patched.append(f"{mod_name}.{attr}")
    for (mod_name, cls_name), names in _METHOD_PATCHES.items():
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:
                continue
        cls = getattr(mod, cls_name, None) if mod is not None else None
        if cls is None:
            continue
        for attr, fn in names.items():
            if attr in cls.__dict__:
                _originals.setdefault((mod_name, cls_name, attr), cls.__dict__[attr])
                setattr(cls, attr, staticmethod(fn))
                patched.append(f"{mod_name}.{cls_name}.{attr}")
    return patched


def uninstall() -> None:
    for key, obj in _originals.items():
        mod = sys.modules.get(key[0])
        if mod is None:
            continue
        if len(key) == 3:
            setattr(getattr(mod, key[1]), key[2], obj)
        else:
            setattr(mod, key[1], obj)
    _originals.clear()
