"""The reference's own CALLERS run unchanged on the drop-in (SURVEY.md section 4 / section 7 hard part 6).

`oracle/build_ref.py::copy_callers` copies the reference's Python callers of the path verbatim into the git-ignored
baseline/_ref/refpy/ (gaussian_renderer/__init__.py, scene/cameras.py, scene/gaussian_model.py, utils/*, and the reference's own
autograd wrapper diff_gaussian_rasterization/__init__.py). This test imports them as they are -- stubbing only the third-party
modules that are absent from the image (tinycudann, plyfile) -- and drives

    GaussianModel.create_from_pcd  (-> simple_knn._C.distCUDA2)        scene/gaussian_model.py:133-163
    gaussian_renderer.render()                                          gaussian_renderer/__init__.py:203-391
    the optimisation loop of train.py:94-186 (loss, backward, densification statistics, densify/prune, Adam)

twice in one process: arm A resolves `diff_gaussian_rasterization` / `simple_knn` to THIS repo's drop-ins, arm B to the
reference's wrapper over its own CUDA build (oracle/_ref/ref_dgr_C.so, ref_knn_C.so). Everything render() returns, every
parameter .grad, viewspace_points.grad and the parameters after the loop are compared. Skips when the files are absent."""
import importlib
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

REFPY = os.path.join(H.ROOT, "baseline", "_ref", "refpy")
CALLER_MODULES = ["gaussian_renderer", "scene", "scene.cameras", "scene.gaussian_model", "utils", "utils.graphics_utils", "utils.general_utils",
                  "utils.sh_utils", "utils.system_utils", "utils.loss_utils", "diff_gaussian_rasterization", "diff_gaussian_rasterization._C",
                  "simple_knn", "simple_knn._C"]


def _have():
    return os.path.exists(os.path.join(REFPY, "gaussian_renderer", "__init__.py")) and H.ref_dgr() is not None and H.ref_knn() is not None


def _purge():
    for m in list(sys.modules):
        if m in CALLER_MODULES or m.startswith(("gaussian_renderer.", "scene.", "utils.")):
            del sys.modules[m]


class _Arm:
    """Context: the reference's callers importable, with `diff_gaussian_rasterization` / `simple_knn` bound to one implementation."""

    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.saved = {m: sys.modules.get(m) for m in CALLER_MODULES + ["tinycudann", "plyfile"]}
        self.path = list(sys.path)
        _purge()
        # third-party modules the callers import at module level but never touch on this path
        sys.modules["tinycudann"] = types.ModuleType("tinycudann")
        ply = types.ModuleType("plyfile")
        ply.PlyData = ply.PlyElement = type("Stub", (), {})
        sys.modules["plyfile"] = ply
        if self.which == "ours":
            Pk = H.pkg()
            sys.modules["diff_gaussian_rasterization"] = Pk.diff_gaussian_rasterization
            sys.modules["simple_knn"] = importlib.import_module("simple_knn")
            sys.modules["simple_knn._C"] = importlib.import_module("simple_knn._C")
        else:  # the reference's own Python wrapper over the reference's own compiled extension
            pkgmod = types.ModuleType("diff_gaussian_rasterization")
            pkgmod.__path__ = [os.path.join(REFPY, "ref_wrapper", "diff_gaussian_rasterization")]
            pkgmod.__package__ = "diff_gaussian_rasterization"
            sys.modules["diff_gaussian_rasterization"] = pkgmod
            sys.modules["diff_gaussian_rasterization._C"] = H.ref_dgr()
            src = os.path.join(REFPY, "ref_wrapper", "diff_gaussian_rasterization", "__init__.py")
            exec(compile(open(src).read(), src, "exec"), pkgmod.__dict__)  # runs `from . import _C` against the module above
            knn = types.ModuleType("simple_knn")
            knn.__path__ = []
            sys.modules["simple_knn"] = knn
            sys.modules["simple_knn._C"] = H.ref_knn()
        sys.path.insert(0, REFPY)
        self.render_mod = importlib.import_module("gaussian_renderer")
        self.model_mod = importlib.import_module("scene.gaussian_model")
        self.cam_mod = importlib.import_module("scene.cameras")
        self.loss_mod = importlib.import_module("utils.loss_utils")
        self.gfx = importlib.import_module("utils.graphics_utils")
        return self

    def __exit__(self, *exc):
        _purge()
        sys.path[:] = self.path
        for m, v in self.saved.items():
            if v is not None:
                sys.modules[m] = v
            else:
                sys.modules.pop(m, None)
        return False


class _Pipe:
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False


class _Opt:  # arguments/__init__.py:90-112, with the schedule compressed so that a short loop reaches densification
    iterations = 30_000
    position_lr_init = 0.00016
    position_lr_final = 0.0000016
    position_lr_delay_mult = 0.01
    position_lr_max_steps = 30_000
    feature_lr = 0.0025
    opacity_lr = 0.05
    segment_lr = 0.05
    scaling_lr = 0.005
    rotation_lr = 0.001
    percent_dense = 0.01
    lambda_dssim = 0.2
    lambda_depth = 0.1
    densification_interval = 3
    opacity_reset_interval = 3000
    densify_from_iter = 2
    densify_until_iter = 15_000
    densify_grad_threshold = 0.0002


def _scene(seed=0, n=6000):
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(n, 3, generator=g) * 2.0 - 1.0).numpy().astype(np.float64)
    cols = torch.rand(n, 3, generator=g).numpy().astype(np.float64)
    W, Hh = 200, 152
    gt = torch.rand(3, Hh, W, generator=g)
    gt_depth = torch.rand(1, Hh, W, generator=g)
    return pts, cols, W, Hh, gt, gt_depth


def _make(arm, pts, cols, W, Hh, gt, gt_depth):
    pcd = arm.gfx.BasicPointCloud(points=pts, colors=cols, normals=np.zeros_like(pts))
    gm = arm.model_mod.GaussianModel(3, num_class=2)
    gm.create_from_pcd(pcd, spatial_lr_scale=1.0)  # simple_knn._C.distCUDA2 inside
    fovx = math.radians(60.0)
    fovy = 2.0 * math.atan(math.tan(fovx / 2.0) * Hh / W)
    cam = arm.cam_mod.Camera(colmap_id=0, R=np.eye(3), T=np.array([0.0, 0.0, 3.0]), FoVx=fovx, FoVy=fovy, image=gt, gt_alpha_mask=None, image_name="v0",
                             uid=0, gt_depth=gt_depth)
    return gm, cam


def _run_render_once(arm, scene, bbox):
    pts, cols, W, Hh, gt, gt_depth = scene
    gm, cam = _make(arm, pts, cols, W, Hh, gt, gt_depth)
    gm.active_sh_degree = 3
    with torch.no_grad():  # give every SH band and the segments something to render
        g = torch.Generator(device="cuda").manual_seed(5)
        gm._features_rest.add_(0.05 * torch.randn(gm._features_rest.shape, device="cuda", generator=g))
        gm._segment.add_(torch.randn(gm._segment.shape, device="cuda", generator=g))
        gm._scaling.add_(0.3 * torch.randn(gm._scaling.shape, device="cuda", generator=g))
        gm._rotation.add_(0.3 * torch.randn(gm._rotation.shape, device="cuda", generator=g))
        gm._opacity.add_(2.0 + torch.randn(gm._opacity.shape, device="cuda", generator=g))
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    mask = None
    if bbox:
        mask = (gm.get_xyz[:, 0] > -0.5).detach()
    pkg = arm.render_mod.render(cam, gm, _Pipe(), bg, bbox_mask=mask)
    loss = (pkg["render"] * torch.linspace(0.5, 1.5, W, device="cuda")).sum() / (W * Hh) + pkg["depth"].mean() + 0.3 * pkg["alpha"].mean() + \
        (pkg["segment"] * torch.linspace(1.0, -1.0, Hh, device="cuda")[None, :, None]).sum() / (W * Hh)
    loss.backward()
    out = {k: pkg[k].detach().clone() for k in ["render", "visibility_filter", "radii", "depth", "alpha", "segment"]}
    out["viewspace_points.grad"] = pkg["viewspace_points"].grad.clone() if pkg["viewspace_points"].grad is not None else None
    for name in ["_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "_segment"]:
        out["grad" + name] = getattr(gm, name).grad.clone()
    out["scales_init"] = gm._scaling.detach().clone()
    return out


NAMES = ["_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "_segment"]


def _train_loop(arm, scene, iters=5):
    """train.py:94-186 with its own modules: render -> L1 + SSIM (+ depth, localrf choice) -> backward -> densification stats ->
    densify/prune -> Adam step."""
    pts, cols, W, Hh, gt, gt_depth = scene
    torch.manual_seed(1234)
    gm, cam = _make(arm, pts, cols, W, Hh, gt, gt_depth)
    opt = _Opt()
    gm.training_setup(opt)
    bg = torch.tensor([0.0, 0.0, 0.0], device="cuda")
    L = arm.loss_mod
    losses, counts = [], []
    for iteration in range(1, iters + 1):
        gm.update_learning_rate(iteration)
        if iteration % 2 == 0:
            gm.oneupSHdegree()
        pkg = arm.render_mod.render(cam, gm, _Pipe(), bg)
        image, vsp, vis, radii, depth = pkg["render"], pkg["viewspace_points"], pkg["visibility_filter"], pkg["radii"], pkg["depth"]
        gt_image = cam.original_image.cuda()
        Ll1 = L.l1_loss(image, gt_image)
        loss = (1.0 - opt.lambda_dssim) * Ll1 + opt.lambda_dssim * (1.0 - L.ssim(image, gt_image))
        loss = loss + L.compute_depth_loss(1 / depth.clamp(1e-6), cam.depth.cuda(), opt.lambda_depth)
        loss.backward()
        with torch.no_grad():
            losses.append(float(loss))
            gm.max_radii2D[vis] = torch.max(gm.max_radii2D[vis], radii[vis])
            gm.add_densification_stats(vsp, vis)
            if iteration > opt.densify_from_iter and iteration % opt.densification_interval == 0:
                gm.densify_and_prune(opt.densify_grad_threshold, 0.005, 2.0, None)
            gm.optimizer.step()
            gm.optimizer.zero_grad(set_to_none=True)
            counts.append(int(gm.get_xyz.shape[0]))
            if iteration == 2:  # the last step before the first densification
                early = {n: getattr(gm, n).detach().clone() for n in NAMES}
    params = {n: getattr(gm, n).detach().clone() for n in NAMES}
    return losses, counts, early, params


def _close(a, b, rel, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert H.rel_linf(a, b) <= rel, (what, H.rel_linf(a, b))


@pytest.mark.parametrize("bbox", [False, True])
def test_reference_render_runs_unchanged_on_the_dropin(bbox):
    if not _have():
        pytest.skip("baseline/_ref/refpy or oracle/_ref not present (run oracle/build_ref.py where /root/reference exists)")
    scene = _scene()
    with _Arm("ours") as arm:
        ours = _run_render_once(arm, scene, bbox)
    with _Arm("reference") as arm:
        ref = _run_render_once(arm, scene, bbox)
    assert torch.equal(ours["scales_init"], ref["scales_init"])  # distCUDA2 through create_from_pcd: bit-exact
    assert torch.equal(ours["radii"], ref["radii"]) and torch.equal(ours["visibility_filter"], ref["visibility_filter"])
    assert int(ours["visibility_filter"].sum()) > 1000
    for k in ["render", "depth", "alpha", "segment"]:
        assert float((ours[k] - ref[k]).abs().max()) <= 1e-5, k
    for k in ours:
        if k.startswith("grad") or k == "viewspace_points.grad":
            assert (ours[k] is None) == (ref[k] is None), k
            _close(ours[k], ref[k], 1e-4, k)


def test_reference_training_loop_runs_unchanged_on_the_dropin():
    if not _have():
        pytest.skip("baseline/_ref/refpy or oracle/_ref not present (run oracle/build_ref.py where /root/reference exists)")
    scene = _scene(seed=3, n=4000)
    with _Arm("ours") as arm:
        lo, co, eo, po = _train_loop(arm, scene)
    with _Arm("reference") as arm:
        lr, cr, er, pr = _train_loop(arm, scene)
    # Both implementations sum gradients with float atomics, so a Gaussian whose accumulated screen-space gradient sits within
    # rounding of the densification threshold may be cloned / split by one arm and not by the other: the counts must agree to a
    # handful, the loss curves to 2e-5, and -- whenever the two arms made identical decisions -- the parameters themselves.
    assert co[-1] != co[0], (co, cr)  # densification did happen
    assert all(abs(a - b) <= 3 for a, b in zip(co, cr)), (co, cr)
    for a, b in zip(lo, lr):
        assert abs(a - b) <= 5e-5 * max(1.0, abs(b)), (lo, lr)
    # Parameters after the two steps before the first densification. Adam with eps = 1e-15 turns the SIGN of a rounding-noise
    # gradient (the rotation gradient of an isotropic Gaussian is analytically zero) into a full +-lr step, so two correct
    # implementations may differ by 2 * lr per step in such entries: compare to the step size, not to an ulp.
    lrs = {"_xyz": _Opt.position_lr_init, "_features_dc": _Opt.feature_lr, "_features_rest": _Opt.feature_lr / 20.0, "_scaling": _Opt.scaling_lr,
           "_rotation": _Opt.rotation_lr, "_opacity": _Opt.opacity_lr, "_segment": _Opt.segment_lr}
    for n in NAMES:
        d = (eo[n] - er[n]).abs()
        assert float(d.max()) <= 2 * 2 * lrs[n] * 1.05, (n, float(d.max()))
        if n in ("_xyz", "_features_dc", "_opacity", "_scaling"):  # parameters whose gradient is signal: the typical entry agrees far better
            assert float(d.quantile(0.5)) <= 0.05 * lrs[n], (n, float(d.quantile(0.5)))
    # after a densification the rows themselves may differ (one arm splits Gaussian i, the other its neighbour j), so the final
    # parameters are compared through the losses above only
    assert all(torch.isfinite(po[n]).all() for n in NAMES)
