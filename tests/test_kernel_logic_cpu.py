"""CPU-side check of the CUDA kernel bodies (tile/halo logic, reflection adjoint, pyramid
tables, closed-form backward) through the host emulator in tests/emu, against the golden
fixtures produced by the reference.  The GPU tier (test_gpu_*.py) repeats these checks through
the real C ABI on a B200."""
import numpy as np
import pytest
import torch

import emu_binding as emu
from helpers import (GOLDEN_CASES, Golden, assert_argmin_matches, assert_grad_close,
                     assert_loss_close)
from oracle import photo_oracle as po


def level_tables(g):
    k = g.z["intrinsics"]
    return np.stack([po.scaled_intrinsics(k, (g.width, g.height), (g.width >> s, g.height >> s))
                     for s in range(g.num_scales)])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_photo_kernel_matches_reference(name):
    g = Golden(name)
    inp = g.inputs()
    out = emu.photo(level_tables(g), inp["images"], inp["depth"], inp["poses"], inp["noise"], g.num_scales)
    assert_loss_close(out["recon"], g.z["ref64_recon"], "recon")
    for s in range(g.num_scales):
        assert_argmin_matches(out["argmin"][s], g, s, "emulated kernel")
    assert_grad_close(out["grad_depth"], g.z["ref64_grad_depth"], "dL/d depth")
    assert_grad_close(out["grad_pose"][0], g.z["ref64_grad_pose0"], "dL/dT0")
    assert_grad_close(out["grad_pose"][1], g.z["ref64_grad_pose1"], "dL/dT1")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_forward_only_matches(name):
    g = Golden(name)
    inp = g.inputs()
    out = emu.photo(level_tables(g), inp["images"], inp["depth"], inp["poses"], inp["noise"], g.num_scales,
                    with_grad=False)
    assert_loss_close(out["recon"], g.z["ref64_recon"], "recon (no grad)")
    for s in range(g.num_scales):
        assert_argmin_matches(out["argmin"][s], g, s, "emulated kernel (no grad)")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_smoothness_matches_reference(name):
    g = Golden(name)
    inp = g.inputs()
    out = emu.smooth(inp["images"][0], inp["disp"])
    assert_loss_close(out["smooth"], g.z["ref64_smooth"], "smooth")
    assert_grad_close(out["grad_disp"], g.z["ref64_grad_disp"], "dL/d disp")


def test_emulated_grad_scales_with_upstream():
    g = Golden("city_wide")
    inp = g.inputs()
    a = emu.photo(level_tables(g), inp["images"], inp["depth"], inp["poses"], inp["noise"], g.num_scales)
    b = emu.photo(level_tables(g), inp["images"], inp["depth"], inp["poses"], inp["noise"], g.num_scales,
                  grad_loss=10.0)
    torch.testing.assert_close(b["grad_depth"], 10.0 * a["grad_depth"], rtol=1e-6, atol=0)
    torch.testing.assert_close(b["grad_pose"][0], 10.0 * a["grad_pose"][0], rtol=1e-6, atol=0)


def test_emulated_standalone_operators():
    g = Golden("city_near")
    inp = g.inputs()
    k = inp["intrinsics"]
    grid = emu.warp_grid(inp["depth"], inp["poses"][1], k)
    np.testing.assert_allclose(grid.numpy(), g.z["op_grid"], rtol=0, atol=2e-6)
    warped = emu.warp_image(inp["images"][2], inp["depth"], inp["poses"][1], k)
    np.testing.assert_allclose(warped.numpy(), g.z["op_warped"], rtol=0, atol=1e-4)  # fp32 coordinate ulp x image slope
    nearest = emu.warp_image(inp["images"][2], inp["depth"], inp["poses"][1], k, mode=1)
    assert (nearest.numpy() != g.z["op_nearest"]).mean() < 1e-3
    up = g.t("op_upstream")
    gd, gp, _ = emu.warp_image_bwd(up, inp["images"][2], inp["depth"], inp["poses"][1], k)
    assert_grad_close(gd, g.z["op_warp_grad_depth"], "warp dL/d depth")
    assert_grad_close(gp, g.z["op_warp_grad_pose"], "warp dL/dT")
    ssim = emu.ssim(inp["images"][1], inp["images"][0])
    # centred statistics are closer to the fp64 value than the reference's own fp32 run is
    ssim64 = po.ssim_loss_map(inp["images"][1].double(), inp["images"][0].double())
    np.testing.assert_allclose(ssim.numpy(), ssim64.numpy(), rtol=0, atol=5e-6)
    np.testing.assert_allclose(ssim.numpy(), g.z["op_ssim"], rtol=0, atol=5e-4)
    gx, gy = emu.ssim_bwd(up, inp["images"][1], inp["images"][0])
    x64 = inp["images"][1].double().requires_grad_(True)
    y64 = inp["images"][0].double().requires_grad_(True)
    (po.ssim_loss_map(x64, y64) * up.cpu().double()).sum().backward()
    assert_grad_close(gx, x64.grad, "ssim dL/dx vs fp64")
    assert_grad_close(gy, y64.grad, "ssim dL/dy vs fp64")
    assert_grad_close(gx, g.z["op_ssim_grad_x"], "ssim dL/dx vs fp32 reference", rtol=1e-3)
    assert_grad_close(gy, g.z["op_ssim_grad_y"], "ssim dL/dy vs fp32 reference", rtol=1e-3)


def test_emulated_object_motion_warp():
    """object_motion_map branch of CoordinateWarper (misc/image_warper.py:133-134)."""
    g = Golden("city_near")
    inp = g.inputs()
    k = inp["intrinsics"]
    gen = torch.Generator().manual_seed(5)
    motion = 0.01 * torch.randn(inp["depth"].shape[0], 3, g.height, g.width, generator=gen)
    depth = inp["depth"].clone().requires_grad_(True)
    pose = inp["poses"][0].clone().requires_grad_(True)
    mo = motion.clone().requires_grad_(True)
    want = po.warp_image(inp["images"][1], depth, pose, k, motion=mo)
    up = torch.randn(want.shape, generator=gen)
    (want * up).sum().backward()
    got = emu.warp_image(inp["images"][1], inp["depth"], inp["poses"][0], k, motion=motion)
    np.testing.assert_allclose(got.numpy(), want.detach().numpy(), rtol=0, atol=1e-4)
    gd, gp, gm = emu.warp_image_bwd(up, inp["images"][1], inp["depth"], inp["poses"][0], k, motion=motion)
    assert_grad_close(gd, depth.grad, "motion warp dL/d depth")
    assert_grad_close(gp, pose.grad, "motion warp dL/dT")
    assert_grad_close(gm, mo.grad, "motion warp dL/d motion")


@pytest.mark.parametrize("w,h,scales", [(32, 32, 5), (40, 18, 4), (34, 66, 3), (64, 64, 6), (48, 40, 1)])
def test_emulated_tiny_and_ragged_sizes(w, h, scales):
    """Coarsest level down to 2x2 (reflection padding of a 2-pixel axis), sizes that are not
    multiples of the tile, single-tile images."""
    from codeps_b200 import synthetic
    import codeps_b200
    from helpers import check_photo_grads
    tb = synthetic.make_batch(2, w, h, (0.9 * w, 0.95 * w, 0.5 * w, 0.5 * h), seed=31, shift_px=1, flip_every_other=True)
    noise = po.draw_noise(2, w, h, scales, seed=8)
    k = codeps_b200.ReconstructionLoss(w, h, None, scales, "cpu")._level_intrinsics(tb.camera_models())
    out = emu.photo(k, tb.images, tb.depth, tb.poses, noise, scales)
    ref = po.loss_and_grads(tb.intrinsics.numpy(), tb.images, tb.depth, tb.disp, tb.poses, noise, scales,
                            dtype=torch.float64, level_intrinsics=list(k))
    assert_loss_close(out["recon"], ref["recon"], "recon")
    for s in range(scales):
        top2 = torch.sort(ref["candidates"][s], dim=1).values[:, :2]
        decided = (top2[:, 1] - top2[:, 0]) > 1e-6
        assert not ((out["argmin"][s] != ref["argmin"][s]) & decided).any()
    inp = dict(intrinsics=tb.intrinsics.numpy(), images=tb.images, depth=tb.depth, disp=tb.disp, poses=tb.poses, noise=noise)
    check_photo_grads(out, inp, scales, f"{w}x{h}", level_intrinsics=list(k), max_masked_frac=0.05)
    sm = emu.smooth(tb.images[0], tb.disp)
    assert_loss_close(sm["smooth"], ref["smooth"], "smooth")


def test_emulated_fused_loss_with_object_motion():
    """object_motion_maps branch of ReconstructionLoss (algos/depth.py:296-303) inside the fused
    kernel: loss, selection and all gradients incl. dL/d motion against the fp64 oracle."""
    from helpers import assert_grad_close_masked, unstable_depth_mask
    g = Golden("kitti_odd")
    inp = g.inputs()
    gen = torch.Generator().manual_seed(12)
    b = inp["depth"].shape[0]
    motions = [0.01 * torch.randn(b, 3, g.height, g.width, generator=gen) for _ in range(2)]
    out = emu.photo(level_tables(g), inp["images"], inp["depth"], inp["poses"], inp["noise"], g.num_scales,
                    motions=motions)
    free = po.loss_and_grads(inp["intrinsics"], inp["images"], inp["depth"], inp["disp"], inp["poses"], inp["noise"],
                             g.num_scales, dtype=torch.float64, motions=motions)
    assert_loss_close(out["recon"], free["recon"], "recon with motion")
    for s in range(g.num_scales):
        top2 = torch.sort(free["candidates"][s], dim=1).values[:, :2]
        decided = (top2[:, 1] - top2[:, 0]) > 1e-6
        assert not ((out["argmin"][s] != free["argmin"][s]) & decided).any()
    ref = po.loss_and_grads(inp["intrinsics"], inp["images"], inp["depth"], inp["disp"], inp["poses"], inp["noise"],
                            g.num_scales, dtype=torch.float64, motions=motions, forced_argmin=out["argmin"])
    mask = unstable_depth_mask(ref, out["argmin"], g.height, g.width).unsqueeze(1)
    assert_grad_close_masked(out["grad_depth"], ref["grad_depth"], mask, "dL/d depth")
    for k in range(2):
        assert_grad_close_masked(out["grad_motion"][k], ref["grad_motion"][k], mask.expand(-1, 3, -1, -1), f"dL/d motion{k}")
        assert_grad_close(out["grad_pose"][k], ref["grad_pose"][k], f"dL/dT{k}", rtol=1e-3)


@pytest.mark.parametrize("w,h,scales", [(96, 48, 5), (132, 70, 4)])
def test_emulated_device_intrinsics_path_is_identical(w, h, scales):
    """Intrinsics read from device memory and rescaled per level in the kernel
    give bit-identical results to the host-scaled per-level values of
    CameraModel.get_scaled_model_image_size -- also for a non-power-of-two pyramid (132x70 -> 16x8)."""
    from codeps_b200 import synthetic
    import codeps_b200
    tb = synthetic.make_batch(2, w, h, (1.1 * w, 1.07 * w, 0.52 * w, 0.47 * h), seed=5, flip_every_other=True)
    noise = po.draw_noise(2, w, h, scales, seed=3)
    k = codeps_b200.ReconstructionLoss(w, h, None, scales, "cpu")._level_intrinsics(tb.camera_models())
    host = emu.photo(k, tb.images, tb.depth, tb.poses, noise, scales)
    dev = emu.photo(None, tb.images, tb.depth, tb.poses, noise, scales, full_res_intrinsics=tb.intrinsics.numpy())
    assert float(host["recon"]) == float(dev["recon"])
    for s in range(scales):
        assert torch.equal(host["argmin"][s], dev["argmin"][s])
    assert torch.equal(host["grad_depth"], dev["grad_depth"])
    assert torch.equal(host["grad_pose"][0], dev["grad_pose"][0]) and torch.equal(host["grad_pose"][1], dev["grad_pose"][1])


@pytest.mark.parametrize("case", ["behind_camera", "general_last_row"])
def test_emulated_degenerate_and_general_poses(case):
    """The literal-formula branch of the warp (cdp_warp_point, `regular == false`):
    points that land behind the camera after the transform (z clamped to 1e-5, zero gradient
    through the clamp, misc/image_warper.py:32) and a general 4x4 matrix whose last row is not
    (0,0,0,1) (homogeneous divide by Q_w, misc/image_warper.py:137-138), as a badly initialised
    pose network produces them."""
    from helpers import check_photo_grads
    g = Golden("city_near")
    inp = g.inputs()
    poses = [p.clone() for p in inp["poses"]]
    if case == "behind_camera":
        poses[0][:, 2, 3] -= 1.4 * float(inp["depth"].median())   # t_z pushes the nearer half behind the camera
        poses[1][0, 2, 3] -= 0.9 * float(inp["depth"].median())
    else:
        poses[0][:, 3, :] = torch.tensor([0.02, -0.03, 0.05, 1.1])
        poses[1][:, 3, :] = torch.tensor([-0.01, 0.015, -0.04, 0.93])
    inp = dict(inp, poses=poses)
    out = emu.photo(level_tables(g), inp["images"], inp["depth"], poses, inp["noise"], g.num_scales)
    free = po.loss_and_grads(inp["intrinsics"], inp["images"], inp["depth"], inp["disp"], poses, inp["noise"],
                             g.num_scales, dtype=torch.float64)
    if case == "behind_camera":  # the case must actually exercise the clamp
        pts = inp["depth"].double() * 1.0 + poses[0][:, 2, 3].double().view(-1, 1, 1, 1)
        assert (pts < 1e-5).float().mean() > 0.2
    assert_loss_close(out["recon"], free["recon"], f"recon {case}")
    for s in range(g.num_scales):
        top2 = torch.sort(free["candidates"][s], dim=1).values[:, :2]
        decided = (top2[:, 1] - top2[:, 0]) > 1e-6
        assert not ((out["argmin"][s] != free["argmin"][s]) & decided).any(), f"level {s}"
    print(check_photo_grads(out, inp, g.num_scales, case, max_masked_frac=0.05))


@pytest.mark.parametrize("w,h,scales,shift_px", [(96, 64, 2, 9), (160, 128, 2, 3), (160, 128, 1, 9)])
def test_emulated_taps_outside_the_staged_source_box(w, h, scales, shift_px):
    """Sample displacements of 7-11 px: the bilinear footprints leave the gather margin of the staged
    source boxes (CDP_SRC_MARGIN = 4) and are served from global memory instead; same results.
    The 160x128 cases contain interior tiles (source boxes inside [1, n-2]: CDP_OPT_INTERIOR fast path
    without reflection / border-clip logic), with footprints inside the boxes (3 px) and leaving them (9 px)."""
    from codeps_b200 import synthetic
    import codeps_b200
    from helpers import check_photo_grads
    tb = synthetic.make_batch(2, w, h, (0.9 * w, 0.95 * w, 0.5 * w, 0.5 * h), seed=17, shift_px=shift_px, flip_every_other=True)
    noise = po.draw_noise(2, w, h, scales, seed=5)
    k = codeps_b200.ReconstructionLoss(w, h, None, scales, "cpu")._level_intrinsics(tb.camera_models())
    out = emu.photo(k, tb.images, tb.depth, tb.poses, noise, scales)
    ref = po.loss_and_grads(tb.intrinsics.numpy(), tb.images, tb.depth, tb.disp, tb.poses, noise, scales,
                            dtype=torch.float64, level_intrinsics=list(k))
    shift = (ref["grids"][0][0][..., 0].double() + 1) / 2 * (w - 1) - torch.arange(w, dtype=torch.float64)
    assert shift_px < 7 or float(shift.abs().max()) > 6.0, "the case must leave the margin"
    assert_loss_close(out["recon"], ref["recon"], "recon")
    for s in range(scales):
        top2 = torch.sort(ref["candidates"][s], dim=1).values[:, :2]
        decided = (top2[:, 1] - top2[:, 0]) > 1e-6
        assert not ((out["argmin"][s] != ref["argmin"][s]) & decided).any()
    inp = dict(intrinsics=tb.intrinsics.numpy(), images=tb.images, depth=tb.depth, disp=tb.disp, poses=tb.poses, noise=noise)
    check_photo_grads(out, inp, scales, "taps outside the source box", level_intrinsics=list(k))


@pytest.mark.parametrize("w,h,scales", [(64, 48, 3), (50, 34, 2), (40, 32, 1)])
def test_emulated_fused_heads_equal_the_separate_conversions(w, h, scales):
    """cdp_photo_args.heads: disparity and 6-DoF parameters in, gradients with
    respect to them out.  Must equal the composition disp_to_depth / transformation_from_parameters
    -> loss -> their backward: bit for bit where the pyramid launch does the conversion
    (W % 4 == 0, even H, >= 2 levels), and through the fallback kernels otherwise."""
    from codeps_b200 import synthetic
    import codeps_b200
    gen = torch.Generator().manual_seed(4)
    tb = synthetic.make_batch(2, w, h, (0.9 * w, 0.95 * w, 0.5 * w, 0.5 * h), seed=12, shift_px=1)
    noise = po.draw_noise(2, w, h, scales, seed=3)
    k = codeps_b200.ReconstructionLoss(w, h, None, scales, "cpu")._level_intrinsics(tb.camera_models())
    aa = [1e-2 * torch.randn(2, 3, generator=gen) for _ in range(2)]
    tr = [torch.tensor([[-1.0 / (0.9 * w), 0.0, 0.0]]).repeat(2, 1) + 1e-3 * torch.randn(2, 3, generator=gen) for _ in range(2)]
    fused = emu.photo_from_heads(k, tb.images, tb.disp, ((aa[0], tr[0]), (aa[1], tr[1])), noise, scales, grad_loss=2.5)
    # the same through the stand-alone conversions
    lo, span = 1.0 / 100.0, 1.0 / 0.1 - 1.0 / 100.0
    depth = (1.0 / (np.float32(lo) + np.float32(span) * tb.disp)).float()
    poses = [emu.pose(aa[0], tr[0], True), emu.pose(aa[1], tr[1], False)]
    for got, want in zip(fused["poses"], poses):
        assert torch.equal(got, want)
    np.testing.assert_allclose(fused["depth"].numpy(), depth.numpy(), rtol=2e-7, atol=0)
    sep = emu.photo(k, tb.images, fused["depth"], poses, noise, scales, grad_loss=2.5)
    assert torch.equal(fused["recon"], sep["recon"])
    assert all(torch.equal(a, b) for a, b in zip(fused["argmin"], sep["argmin"]))
    want_gdisp = -(sep["grad_depth"] * np.float32(span) * fused["depth"] * fused["depth"])
    np.testing.assert_allclose(fused["grad_disp"].numpy(), want_gdisp.numpy(), rtol=1e-6, atol=1e-12)
    for kk in range(2):
        _, ga, gt = emu.pose(aa[kk], tr[kk], kk == 0, grad_out=sep["grad_pose"][kk])
        np.testing.assert_allclose(fused["grad_pose_params"][2 * kk].numpy(), ga.numpy(), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(fused["grad_pose_params"][2 * kk + 1].numpy(), gt.numpy(), rtol=1e-5, atol=1e-9)
    # and against the oracle: loss of the reference's own op sequence on the same head outputs
    ref_depth = po.disp_to_depth(tb.disp.double())
    ref_poses = [po.transformation_from_parameters(aa[0].double().view(2, 1, 3), tr[0].double().view(2, 1, 3), True),
                 po.transformation_from_parameters(aa[1].double().view(2, 1, 3), tr[1].double().view(2, 1, 3), False)]
    want = po.reconstruction_loss(tb.intrinsics.numpy(), [i.double() for i in tb.images], ref_depth, ref_poses, noise, scales,
                                  level_intrinsics=list(k))
    assert_loss_close(fused["recon"], want, "recon from heads", rtol=2e-5)


@pytest.mark.parametrize("hw,levels", [((512, 1024), 5), ((376, 1408), 5), ((384, 1280), 5), ((33, 65), 2), ((32, 32), 1),
                                       ((2048, 4096), 5), ((96, 4000), 3), ((1000, 40), 4)])
def test_tile_lookup_matches_integer_arithmetic(hw, levels):
    """cdp_tile_ctx finds a block's level without a loop and its tile row with the host's reciprocal
    (a multiply-high instead of a division, on the critical path to the tile's TMA loads): every
    block must land on the tile plain integer arithmetic gives, incl. one-tile-wide levels."""
    import emu_binding
    h, w = hw
    got = emu_binding.tile_table(h, w, levels)
    want = []
    for s in range(levels):
        hs, ws = h >> s, w >> s
        tiles_x, tiles_y = (ws + 31) // 32, (hs + 31) // 32
        for tile in range(tiles_x * tiles_y):
            want.append((s, (tile % tiles_x) * 32, (tile // tiles_x) * 32))
    want = np.asarray(want, dtype=np.int32)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)
