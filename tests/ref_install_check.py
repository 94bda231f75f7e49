"""Runs in a subprocess of tests/test_reference_integration.py (it imports the real CoDEPS checkout
and rebinds names in it, which must not leak into the pytest process).  Prints one JSON report."""
import importlib
import inspect
import json
import os
import sys
import types

REFERENCE = sys.argv[1]
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REFERENCE)
sys.path.insert(0, REPO)

# the three optional dependencies of the checkout that are not on the hot path
yc = types.ModuleType("yacs.config")
yc.CfgNode = type("CfgNode", (dict,), {})
sys.modules["yacs"] = types.ModuleType("yacs")
sys.modules["yacs.config"] = yc
ex = types.ModuleType("skimage.exposure")
ex.match_histograms = ex.is_low_contrast = None
sys.modules["skimage"] = types.ModuleType("skimage")
sys.modules["skimage.exposure"] = ex
kc = types.ModuleType("kornia.contrib")
kc.connected_components = kc.distance_transform = None
sys.modules["kornia"] = types.ModuleType("kornia")
sys.modules["kornia.contrib"] = kc

import torch  # noqa: E402

report = {"signatures": {}, "errors": []}


def sig(fn):
    """Parameter names, kinds and defaults (annotations are not compared: ours are looser)."""
    out = []
    for name, prm in inspect.signature(fn).parameters.items():
        default = None if prm.default is inspect.Parameter.empty else repr(prm.default)
        out.append((name, str(prm.kind), default))
    return out


# 1. originals, before install
CLASSES = [("misc.camera_model", "CameraModel"), ("misc.image_warper", "ImageWarper"),
           ("misc.image_warper", "CoordinateWarper"), ("algos.depth", "SSIMLoss"),
           ("algos.depth", "ReconstructionLoss"), ("algos.depth", "EdgeAwareSmoothnessLoss"),
           ("algos.depth", "FlowSmoothnessLoss"), ("algos.depth", "FlowSparsityLoss"), ("eval.depth", "DepthEvaluator")]
STATIC = [("models.pose_head", "PoseHead", "transformation_from_parameters"),
          ("models.depth_head", "DepthHead", "disp_to_depth"), ("datasets.mixup", "Mixup", "warp_c2c")]
originals = {}
for mod, name in CLASSES:
    cls = getattr(importlib.import_module(mod), name)
    entry = {"init": sig(cls.__init__)}
    call = "forward" if "forward" in cls.__dict__ else "__call__"
    entry["call_name"] = call
    entry["call"] = sig(getattr(cls, call))
    entry["methods"] = sorted(n for n, v in cls.__dict__.items() if callable(v) and not n.startswith("_"))
    originals[(mod, name)] = entry
for mod, cls_name, fn in STATIC:
    cls = getattr(importlib.import_module(mod), cls_name)
    originals[(mod, cls_name, fn)] = {"call": sig(getattr(cls, fn))}
importlib.import_module("codeps.model_setup")
importlib.import_module("codeps.online_adap")

# 2. install
import codeps_b200  # noqa: E402
report["patched"] = codeps_b200.install()

# 3. every rebound class has the reference's constructor and call signature and public methods
for (mod, name), want in [(k, v) for k, v in originals.items() if len(k) == 2]:
    cls = getattr(sys.modules[mod], name)
    ours = cls.__module__.startswith("codeps_b200")
    got_call = getattr(cls, want["call_name"], None) or getattr(cls, "__call__")
    entry = {"ours": ours, "init_equal": sig(cls.__init__) [:len(want["init"])] == want["init"],
             "extra_init_have_defaults": all(d is not None for _, _, d in sig(cls.__init__)[len(want["init"]):]),
             "call_equal": sig(got_call) == want["call"],
             "missing_methods": [m for m in want["methods"] if not hasattr(cls, m)]}
    if not entry["init_equal"]:
        entry["init_ref"], entry["init_ours"] = want["init"], sig(cls.__init__)
    if not entry["call_equal"]:
        entry["call_ref"], entry["call_ours"] = want["call"], sig(got_call)
    report["signatures"][f"{mod}.{name}"] = entry
for (mod, cls_name, fn), want in [(k, v) for k, v in originals.items() if len(k) == 3]:
    cls = getattr(sys.modules[mod], cls_name)
    got = getattr(cls, fn)
    report["signatures"][f"{mod}.{cls_name}.{fn}"] = {
        "ours": getattr(got, "__module__", "").startswith("codeps_b200"), "init_equal": True,
        "extra_init_have_defaults": True, "call_equal": sig(got) == want["call"], "missing_methods": [],
        "call_ref": want["call"], "call_ours": sig(got)}

# 4. the constructor calls of codeps/model_setup.py:63-85, with the names as that module now sees them
ms = sys.modules["codeps.model_setup"]
device = torch.device("cpu")
try:
    ssim_loss = ms.SSIMLoss()
    rec = {"target": ms.ReconstructionLoss(1408, 384, ssim_loss, 5, device),
           "source": ms.ReconstructionLoss(1024, 512, ssim_loss, 5, device)}
    smth = ms.EdgeAwareSmoothnessLoss()
    flow_smth, flow_sparse = ms.FlowSmoothnessLoss(), ms.FlowSparsityLoss()
    depth_eval = ms.DepthEvaluator(True, [0.1, 80.0])
    algo = ms.DepthAlgo(rec["target"], smth, depth_eval, flow_smth, flow_sparse, rec["source"], None)
    report["depth_algo"] = {
        "is_reference_class": type(algo).__module__ == "algos.depth",
        "loss_classes": [type(algo.reconstruction_loss).__module__, type(algo.reconstruction_loss_adapt_source).__module__,
                         type(algo.smoothness_loss).__module__, type(algo.evaluator).__module__,
                         type(algo.flow_smoothness_loss).__module__, type(algo.flow_sparsity_loss).__module__],
        "training_signature": [p for p in inspect.signature(algo.training).parameters],
        "image_warpers": sorted(rec["target"].image_warpers), "scaled_width": rec["target"].scaled_width[4]}
    cam = sys.modules["codeps.online_adap"].CameraModel.from_tensor(1408, 384, torch.tensor([552.5, 552.5, 682.0, 238.7]))
    report["camera_model"] = {"module": type(cam).__module__, "fx": float(cam.intrinsics["fx"]),
                              "scaled_fx": float(cam.get_scaled_model_image_size(704, 192).intrinsics["fx"])}
except Exception as exc:  # reported, the test asserts on it
    report["errors"].append(repr(exc))

codeps_b200.uninstall()
report["restored"] = all(getattr(sys.modules[m], n).__module__ == m or not getattr(sys.modules[m], n).__module__.startswith("codeps_b200")
                         for m, n in CLASSES)
print("REPORT " + json.dumps(report))
