"""CPU tier: host-side logic, the C-ABI surface, and that nothing falls back to the CPU."""
import ctypes
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

import codeps_b200
from codeps_b200 import _native, build, synthetic
from helpers import Golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build_native()
    assert os.path.exists(path)
    header = open(os.path.join(REPO, "include", "codeps_photo.h")).read()
    declared = set(re.findall(r"\b(cdp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = ctypes.CDLL(path)  # loads without a GPU; no compute call is made here
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/codeps_photo.h but not exported"
    lib.cdp_version.restype = ctypes.c_int
    assert lib.cdp_version() == _native.ABI_VERSION
    sass = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass, sass


def test_host_only_entry_points():
    lib = _native.load()
    n = lib.cdp_resize_tables_bytes(376, 1408, 5)
    assert n > 0 and n % 16 == 0
    buf = np.zeros(n, dtype=np.uint8)
    assert lib.cdp_resize_tables_build(376, 1408, 5, buf.ctypes.data, n) == 0
    assert lib.cdp_resize_tables_build(376, 1408, 5, buf.ctypes.data, 8) == -4  # CDP_ERR_WORKSPACE
    assert b"too small" in lib.cdp_last_error()
    assert lib.cdp_resize_tables_bytes(16, 16, 5) == 0  # level 4 would be 1x1: invalid
    assert lib.cdp_photo_scratch_bytes(8, 512, 1024, 5, 0) > 0
    assert lib.cdp_photo_scratch_bytes(8, 512, 1024, 5, 1) > lib.cdp_photo_scratch_bytes(8, 512, 1024, 5, 0)
    assert lib.cdp_photo_saved_bytes(8, 512, 1024, 5, 1) > lib.cdp_photo_saved_bytes(8, 512, 1024, 5, 0)
    assert lib.cdp_photo_scratch_bytes(8, 512, 1024, 9, 0) == 0
    # taps for ratio 2: every output reads two neighbouring inputs with weight 1/2
    rec = buf.view(np.int32).reshape(-1, 4)
    first = rec[0]
    assert first[0] == 0 and first[1] == 1
    assert np.frombuffer(first[2:].tobytes(), dtype=np.float32).tolist() == [0.5, 0.5]


def test_resize_tables_match_torch_interpolate():
    """The tap tables reproduce F.interpolate(bilinear, align_corners=False) for integer and
    non-integer ratios (checked through the emulator's table builder = same code)."""
    import emu_binding as emu
    import torch.nn.functional as F
    for h, w, levels in ((376, 1408, 5), (70, 132, 5), (64, 64, 3)):
        tab = emu.resize_tables(h, w, levels).numpy()
        lib = _native.load()
        n = lib.cdp_resize_tables_bytes(h, w, levels)
        prod = np.zeros(n, dtype=np.uint8)
        assert lib.cdp_resize_tables_build(h, w, levels, prod.ctypes.data, n) == 0
        assert np.array_equal(tab, prod)
        rec_i = tab.view(np.int32).reshape(-1, 4)
        rec_f = tab.view(np.float32).reshape(-1, 4)
        off = 0
        x = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(1, 1, h, w).contiguous()
        for s in range(1, levels):
            ws, hs = w >> s, h >> s
            want = F.interpolate(x, (hs, ws), mode="bilinear", align_corners=False)[0, 0, 0]
            i0, i1, w0, w1 = rec_i[off:off + ws, 0], rec_i[off:off + ws, 1], rec_f[off:off + ws, 2], rec_f[off:off + ws, 3]
            got = i0 * w0 + i1 * w1
            np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=2e-4)
            off += ws + hs + w + h


def test_camera_model_surface():
    cam = codeps_b200.CameraModel(1024, 512, 1131.26, 1132.65, 548.49, 256.5685)
    assert cam.image_size == {"width": 1024, "height": 512}
    assert list(cam.intrinsics) == ["fx", "fy", "cx", "cy"]
    half = cam.get_scaled_model_image_size(512, 256)
    assert half.image_size["width"] == 512 and half.intrinsics["fx"] == 1131.26 * 0.5
    assert cam.get_scaled_model(0.5, 0.25).intrinsics["cy"] == 256.5685 * 0.25
    t = cam.to_tensor()
    back = codeps_b200.CameraModel.from_tensor(1024, 512, t)
    assert isinstance(back.intrinsics["fx"], np.float32) and np.isclose(back.intrinsics["cx"], 548.49)
    u, v = cam.get_image_point(torch.tensor([1.0]), torch.tensor([2.0]), torch.tensor([4.0]))
    assert torch.allclose(u, torch.tensor([1131.26 / 4 + 548.49])) and torch.allclose(v, torch.tensor([1132.65 / 2 + 256.5685]))
    rx, ry, rz = cam.get_viewing_ray(torch.tensor([548.49]), torch.tensor([256.5685]))
    assert torch.allclose(rz, torch.ones(1)) and float(rx) == 0.0 and float(ry) == 0.0
    with pytest.raises(AssertionError):
        codeps_b200.CameraModel(0, 512, 1.0, 1.0, 1.0, 1.0)
    with pytest.raises(AssertionError):
        codeps_b200.CameraModel(10, 10, -1.0, 1.0, 1.0, 1.0)


def test_level_intrinsics_follow_reference_rounding():
    from oracle import photo_oracle as po
    g = Golden("kitti_odd")
    cams = [codeps_b200.CameraModel.from_tensor(g.width, g.height, torch.from_numpy(k)) for k in g.z["intrinsics"]]
    loss = codeps_b200.ReconstructionLoss(g.width, g.height, None, g.num_scales, "cpu")
    got = loss._level_intrinsics(cams)
    for s in range(g.num_scales):
        want = po.scaled_intrinsics(g.z["intrinsics"], (g.width, g.height), (g.width >> s, g.height >> s))
        assert np.array_equal(got[s], want)
    assert loss.scaled_width[4] == g.width // 16 and loss.scaled_height[4] == g.height // 16
    assert isinstance(loss.image_warpers[0], codeps_b200.ImageWarper)


def test_algorithmic_bytes_match_baseline_md():
    assert synthetic.algorithmic_bytes(1024, 512) == 76_107_776
    assert synthetic.algorithmic_bytes(1408, 376) == 76_845_296
    assert synthetic.algorithmic_bytes(1280, 384) == 71_351_040


def test_cpu_tensors_are_rejected_not_computed():
    """There is no CPU path: CPU inputs raise instead of silently running somewhere else."""
    tb = synthetic.make_batch(1, 64, 32, (70.0, 70.0, 32.0, 16.0), seed=1)
    loss = codeps_b200.ReconstructionLoss(64, 32, codeps_b200.SSIMLoss(), 3, "cpu")
    with pytest.raises(RuntimeError, match="CUDA only"):
        loss(tb.camera_models(), tb.images, tb.depth, tb.poses)
    with pytest.raises(RuntimeError, match="CUDA only"):
        codeps_b200.EdgeAwareSmoothnessLoss()(tb.images[0], tb.disp)
    with pytest.raises(RuntimeError, match="CUDA only"):
        codeps_b200.SSIMLoss()(tb.images[0], tb.images[1])
    with pytest.raises(RuntimeError, match="CUDA only"):
        codeps_b200.ImageWarper(64, 32, "cpu")(tb.camera_models(), tb.images[1], tb.depth, tb.poses[0])
    with pytest.raises(NotImplementedError):
        codeps_b200.SSIMLoss(window_size=5)


def test_product_package_never_imports_the_oracle_or_emulator():
    for root, _, files in os.walk(os.path.join(REPO, "codeps_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "photo_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                assert "cdp_emu" not in text and "emu_binding" not in text, f


def test_install_rebinds_reference_names():
    """install() swaps the classes into (stand-ins for) the reference's modules."""
    names = ["misc", "misc.camera_model", "misc.image_warper", "algos", "algos.depth", "codeps", "codeps.model_setup",
             "datasets", "datasets.mixup", "eval", "eval.depth"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names:
            sys.modules[n] = types.ModuleType(n)
        sentinel = object()
        sys.modules["algos.depth"].ReconstructionLoss = sentinel
        sys.modules["algos.depth"].SSIMLoss = sentinel
        sys.modules["algos.depth"].ImageWarper = sentinel
        sys.modules["misc"].ImageWarper = sentinel
        sys.modules["codeps.model_setup"].EdgeAwareSmoothnessLoss = sentinel
        sys.modules["algos.depth"].FlowSparsityLoss = sentinel
        sys.modules["eval.depth"].DepthEvaluator = sentinel

        class Mixup:  # stand-in for datasets.mixup.Mixup with its static method
            @staticmethod
            def warp_c2c(*args, **kwargs):
                return sentinel

            def embed(self):
                return self.warp_c2c()
        sys.modules["datasets.mixup"].Mixup = Mixup
        patched = codeps_b200.install(import_missing=False)
        assert "datasets.mixup.Mixup.warp_c2c" in patched and "algos.depth.FlowSparsityLoss" in patched
        assert sys.modules["eval.depth"].DepthEvaluator is codeps_b200.DepthEvaluator
        assert Mixup.warp_c2c is codeps_b200.warp_c2c and Mixup().warp_c2c is codeps_b200.warp_c2c
        assert "algos.depth.ReconstructionLoss" in patched and "misc.ImageWarper" in patched
        assert sys.modules["algos.depth"].ReconstructionLoss is codeps_b200.ReconstructionLoss
        assert sys.modules["codeps.model_setup"].EdgeAwareSmoothnessLoss is codeps_b200.EdgeAwareSmoothnessLoss
        assert not hasattr(sys.modules["misc"], "CameraModel")  # only existing names are rebound
        codeps_b200.uninstall()
        assert sys.modules["algos.depth"].ReconstructionLoss is sentinel
        assert Mixup().embed() is sentinel
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_synthetic_batches_are_seeded_and_shaped():
    a = synthetic.make_preset_batch("kitti360", 2, seed=4, flip_every_other=True)
    b = synthetic.make_preset_batch("kitti360", 2, seed=4, flip_every_other=True)
    assert torch.equal(a.images[1], b.images[1]) and torch.equal(a.poses[0], b.poses[0])
    assert a.images[0].shape == (2, 3, 376, 1408) and a.depth.shape == (2, 1, 376, 1408)
    assert a.intrinsics[1, 2] == 1408 - a.intrinsics[0, 2] - 1
    assert float(a.depth.min()) >= 0.1 and float(a.depth.max()) <= 100.0
    assert synthetic.level_sizes(1408, 376, 5)[-1] == (88, 23)
