"""ORACLE (test infrastructure, NOT product code) -- dense per-pixel Gaussian-splatting renderer in PyTorch.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this file.  The product path (`log_b200/`) never does.

PARITY STATUS: **partially pinned**.
  * projection / 3D covariance / EWA 2D covariance / radius follow the reference's own PyTorch twin
    `LoG/model/geometry.py:4-41, 91-151` and its CUDA twin `LoG/cuda/compute_radius_kernel.cu:28-156`;
    these are pinned by golden vectors generated from the reference itself (`tests/golden/make_golden.py`).
  * the SH basis follows `LoG/model/sh_utils.py:1-68` (+0.5 / C0 from `activation.py:27-34`); pinned the same way.
  * camera conventions follow `LoG/dataset/base.py:20-55` and `LoG/utils/camera.py:7-29`
    (matrices stored transposed, row-vector convention).
  * the blend itself (tile binning, depth sort, alpha rule, early stop, backward) lives in the un-vendored,
    un-pinned `diff_gaussian_rasterization[_wodilate]` packages (`docs/install.md:37-43`).  Those sources are NOT
    in /root/reference, so this part restates the published 3DGS algorithm (Kerbl et al. 2023) from general
    knowledge and is **parity unpinned** against LoG's binaries.  Constants stated here are THIS repo's definition:
        tile 16x16; alpha = min(0.99, o * exp(power)); power > 0 skipped; alpha < 1/255 skipped;
        stop before the Gaussian that would take T below 1e-4; near cull view-z <= 0.2;
        backward = autograd of this forward with masks constant, the 0.99 clamp straight-through and the
        1.3*tanfov clamp of t.x/t.y treated as a constant (both as in the published backward);
        radius = ceil(3 sqrt(lambda_max)); tile rectangle from the radius square; out = C + T_final * bg.

The oracle is "dense": for every pixel it evaluates every Gaussian (masked), in global (depth, index) order,
with transmittance by cumulative product.  Gradients come from autograd, so the backward is defined by the
forward, not by a second hand derivation.  Works in float64 (default) or float32.
"""
import math
from typing import NamedTuple, Optional

import torch

from log_b200.synthetic import Camera, make_camera, make_scene, make_cotangent  # noqa: F401  (neutral, torch-only)

TILE = 16
NEAR_Z = 0.2
ALPHA_MAX = 0.99
ALPHA_MIN = 1.0 / 255.0
T_STOP = 1e-4
FILTER_VAR = 0.3       # LoG/cuda/compute_radius_kernel.cu:61  (#define DILATE_PIXEL 0.3)
CLAMP_FOV = 1.3        # compute_radius_kernel.cu:71-72

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]

FILTER_ADD = 0    # stock 3DGS: cov.xx += 0.3, cov.yy += 0.3                       [B]
FILTER_MAX = 1    # LoG / "wodilate": cov.xx = max(cov.xx, 0.3) (compute_radius_kernel.cu:100-103)  [V for radius]
FILTER_NONE = 2   # fork with use_filter=False (renderer.py:151-152) -- assumption  [I]


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """geometry.py:4-25 WITHOUT the normalisation: the CUDA kernel does not normalise
    (compute_radius_kernel.cu:36 `glm::vec4 q = rot;// / glm::length(rot);`); LoG pre-normalises (activation.py:41)."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(-1, 3, 3)


def cov3d(scales: torch.Tensor, rotations: torch.Tensor, scale_modifier: float = 1.0) -> torch.Tensor:
    """geometry.py:27-41 : Sigma = (R S)(R S)^T."""
    R = quat_to_rotmat(rotations)
    L = R * (scales * scale_modifier)[:, None, :]
    return L @ L.transpose(-1, -2)


def cov3d_from_precomp(cov6: torch.Tensor) -> torch.Tensor:
    """The stock API's cov3D_precomp (N,6) = upper triangle (xx, xy, xz, yy, yz, zz) of the world-space covariance
    (diff_gaussian_rasterization: computeCov3D's output layout, read back by computeCov2D) -> (N,3,3) symmetric.  The
    stock rasteriser uses it as given: no scale modifier."""
    xx, xy, xz, yy, yz, zz = cov6.unbind(-1)
    return torch.stack([xx, xy, xz, xy, yy, yz, xz, yz, zz], dim=-1).reshape(-1, 3, 3)


def cov2d(Sigma, means3D, cam: Camera, filter_mode: int):
    """geometry.py:91-130 (computeCov2D0) / compute_radius_kernel.cu:63-105.
    Returns (a, b, c) = (cov_xx, cov_xy, cov_yy) after the low-pass filter, and the view-space point t."""
    V = cam.viewmatrix
    t = means3D @ V[:3, :3] + V[3:, :3]
    tx, ty, tz = t[:, 0], t[:, 1], t[:, 2]
    fx = cam.image_width / (2.0 * cam.tanfovx)
    fy = cam.image_height / (2.0 * cam.tanfovy)
    limx, limy = CLAMP_FOV * cam.tanfovx, CLAMP_FOV * cam.tanfovy
    txtz, tytz = tx / tz, ty / tz
    inx = (txtz >= -limx) & (txtz <= limx)
    iny = (tytz >= -limy) & (tytz <= limy)
    # [B] stock backward treats the clamped t.x as a constant (x_grad_mul = 0): detach when clamped.
    txc = torch.where(inx, tx, (txtz.clamp(-limx, limx) * tz).detach())
    tyc = torch.where(iny, ty, (tytz.clamp(-limy, limy) * tz).detach())
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * txc) / (tz * tz),
                     zero, fy / tz, -(fy * tyc) / (tz * tz)], dim=-1).reshape(-1, 2, 3)
    W = V[:3, :3].t()
    T = J @ W                                  # (N,2,3)
    cov = T @ Sigma @ T.transpose(-1, -2)      # (N,2,2)
    a, b, c = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]
    if filter_mode == FILTER_ADD:
        a, c = a + FILTER_VAR, c + FILTER_VAR
    elif filter_mode == FILTER_MAX:
        a, c = a.clamp_min(FILTER_VAR), c.clamp_min(FILTER_VAR)
    return a, b, c, t


def radius_from_cov(a, b, c):
    """compute_radius_kernel.cu:139-152 / geometry.py:141-151 (un-ceiled float radius)."""
    det = a * c - b * b
    mid = 0.5 * (a + c)
    root = torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    lam = torch.maximum(mid + root, mid - root)
    return 3.0 * torch.sqrt(lam), det


def compute_radius(means3D, scales, rotations, cam: Camera):
    """`rasterizer.compute_radius(xyz, scaling, rotation)` (level_of_gaussian.py:59) ==
    LoG/cuda compute_radius_cuda (compute_radius_kernel.cu:107-156):
    NDC cull +-1.3, NO near cull, max(.,0.3) filter, det==0 -> 0, float radius (no ceil)."""
    P = cam.projmatrix
    hom = means3D @ P[:3, :] + P[3:, :]
    pw = 1.0 / (hom[:, 3] + 1e-7)
    px, py = hom[:, 0] * pw, hom[:, 1] * pw
    keep = ~((px < -1.3) | (px > 1.3) | (py < -1.3) | (py > 1.3))
    Sigma = cov3d(scales, rotations, 1.0)
    a, b, c, _ = cov2d(Sigma, means3D, cam, FILTER_MAX)
    rad, det = radius_from_cov(a, b, c)
    keep = keep & (det != 0)
    return torch.where(keep, rad, torch.zeros_like(rad))


def eval_sh(deg: int, shs: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """SH colour WITH the DC term: shs (N,K,3), K >= (deg+1)^2.  Basis == sh_utils.py:31-58 shifted by one
    (their `sh[...,0]` is our shs[:,1]); `+0.5` and C0 from sh_utils.py:69-73.  Clamp at 0 is [B] stock behaviour."""
    res = C0 * shs[:, 0]
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        res = res - C1 * y * shs[:, 1] + C1 * z * shs[:, 2] - C1 * x * shs[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * shs[:, 4] + C2[1] * yz * shs[:, 5] + C2[2] * (2.0 * zz - xx - yy) * shs[:, 6]
                   + C2[3] * xz * shs[:, 7] + C2[4] * (xx - yy) * shs[:, 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * shs[:, 9] + C3[1] * xy * z * shs[:, 10]
                       + C3[2] * y * (4 * zz - xx - yy) * shs[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12]
                       + C3[4] * x * (4 * zz - xx - yy) * shs[:, 13] + C3[5] * z * (xx - yy) * shs[:, 14]
                       + C3[6] * x * (xx - 3 * yy) * shs[:, 15])
    return res + 0.5


def project(means3D, scales, rotations, cam: Camera, filter_mode: int, means2D: Optional[torch.Tensor] = None,
            cov3D_precomp: Optional[torch.Tensor] = None):
    """Per-Gaussian stage.  Returns dict with pixel centre xy (N,2), depth, conic (N,3), float/ceil radius,
    tile rect (N,4) and `valid`.  [B] for cull/rect rules, [V] for the covariance algebra."""
    P = cam.projmatrix
    W_, H_ = cam.image_width, cam.image_height
    hom = means3D @ P[:3, :] + P[3:, :]
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    if means2D is not None:          # dummy zero input whose .grad is dL/d(ndc xy)  (renderer.py:135, counter.py:40)
        ndc = ndc + means2D[:, :2]
    Sigma = cov3d_from_precomp(cov3D_precomp) if cov3D_precomp is not None else cov3d(scales, rotations, cam.scale_modifier)
    a, b, c, t = cov2d(Sigma, means3D, cam, filter_mode)
    depth = t[:, 2]
    radf, det = radius_from_cov(a, b, c)
    valid = (depth > NEAR_Z) & (det > 0)
    det_s = torch.where(valid, det, torch.ones_like(det))
    conic = torch.stack([c / det_s, -b / det_s, a / det_s], dim=-1)
    radius = torch.ceil(radf.detach())
    xy = torch.stack([((ndc[:, 0] + 1.0) * W_ - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H_ - 1.0) * 0.5], dim=-1)
    gx, gy = (W_ + TILE - 1) // TILE, (H_ + TILE - 1) // TILE
    xyd = xy.detach()
    # (int) cast truncates toward zero; after the clamp at 0 that equals floor-then-clamp.
    x0 = torch.trunc((xyd[:, 0] - radius) / TILE).clamp(0, gx)
    x1 = torch.trunc((xyd[:, 0] + radius + TILE - 1) / TILE).clamp(0, gx)
    y0 = torch.trunc((xyd[:, 1] - radius) / TILE).clamp(0, gy)
    y1 = torch.trunc((xyd[:, 1] + radius + TILE - 1) / TILE).clamp(0, gy)
    valid = valid & ((x1 - x0) * (y1 - y0) > 0)
    rect = torch.stack([x0, y0, x1, y1], dim=-1).long()
    rect = torch.where(valid[:, None], rect, torch.zeros_like(rect))
    return dict(xy=xy, depth=depth, conic=conic, cov=(a, b, c), radius=torch.where(valid, radius, torch.zeros_like(radius)),
                radius_f=radf, rect=rect, valid=valid)


def render(means3D, opacities, scales, rotations, cam: Camera, colors_precomp=None, shs=None,
           filter_mode: int = FILTER_ADD, means2D=None, pixel_chunk: int = 8192, return_aux: bool = True, cov3D_precomp=None):
    """Full forward.  Returns dict(image (3,H,W), radii (N,) int, point_id_pixel (H,W) long, point_weight_pixel (H,W),
    point_weight (N,), final_T (H,W), n_instances D).  Differentiable w.r.t. every float input via autograd."""
    N = means3D.shape[0]
    H_, W_ = cam.image_height, cam.image_width
    dt, dev = means3D.dtype, means3D.device
    pr = project(means3D, scales, rotations, cam, filter_mode, means2D, cov3D_precomp)
    if colors_precomp is None:
        dirs = means3D - cam.campos[None]
        dirs = dirs / torch.linalg.norm(dirs, dim=-1, keepdim=True)
        colors = torch.clamp_min(eval_sh(cam.sh_degree, shs, dirs), 0.0)
    else:
        colors = colors_precomp
    C_ = colors.shape[1]
    valid = pr['valid']
    idx = torch.nonzero(valid)[:, 0]
    # global (depth, index) order == per-tile stable sort of index-ordered duplicates
    order = idx[torch.argsort(pr['depth'].detach()[idx], stable=True)]
    xy, conic, rect = pr['xy'][order], pr['conic'][order], pr['rect'][order]
    op = opacities.reshape(-1)[order]
    col = colors[order]
    M = order.shape[0]

    image = []
    finalT = []
    pid = torch.full((H_ * W_,), -1, dtype=torch.long, device=dev)
    pwp = torch.zeros((H_ * W_,), dtype=dt, device=dev)
    pweight_sorted = torch.zeros((M,), dtype=dt, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H_, device=dev), torch.arange(W_, device=dev), indexing='ij')
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    for s in range(0, H_ * W_, pixel_chunk):
        px, py = xs[s:s + pixel_chunk], ys[s:s + pixel_chunk]
        tx, ty = (px // TILE)[:, None], (py // TILE)[:, None]
        in_tile = (tx >= rect[None, :, 0]) & (tx < rect[None, :, 2]) & (ty >= rect[None, :, 1]) & (ty < rect[None, :, 3])
        dx = xy[None, :, 0] - px[:, None].to(dt)
        dy = xy[None, :, 1] - py[:, None].to(dt)
        power = -0.5 * (conic[None, :, 0] * dx * dx + conic[None, :, 2] * dy * dy) - conic[None, :, 1] * dx * dy
        raw = op[None, :] * torch.exp(power)
        # [B] the published backward ignores the 0.99 clamp (dL/dG = opacity * dL/dalpha): straight-through
        alpha = raw + (torch.clamp_max(raw, ALPHA_MAX) - raw).detach()
        keep = in_tile & (power <= 0) & (alpha >= ALPHA_MIN)
        alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
        Tincl = torch.cumprod(1.0 - alpha, dim=1)
        live = Tincl.detach() >= T_STOP            # first failure stops the pixel; Tincl is non-increasing
        alpha = torch.where(live, alpha, torch.zeros_like(alpha))
        Tincl = torch.cumprod(1.0 - alpha, dim=1)
        Texcl = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], dim=1)
        w = alpha * Texcl                          # (P, M)
        Tf = Tincl[:, -1] if M > 0 else torch.ones(px.shape[0], dtype=dt, device=dev)
        image.append(w @ col + Tf[:, None] * cam.bg[None].to(dt))
        finalT.append(Tf)
        if return_aux and M > 0:
            wd = w.detach()
            wmax, amax = wd.max(dim=1)             # first maximum in depth order
            # torch.max returns *an* index of the max; enforce "first" explicitly
            first = (wd == wmax[:, None]).to(torch.int8).argmax(dim=1)
            hit = wmax > 0
            pid[s:s + pixel_chunk] = torch.where(hit, order[first], torch.full_like(first, -1))
            pwp[s:s + pixel_chunk] = wmax
            pweight_sorted = torch.maximum(pweight_sorted, wd.max(dim=0).values)
    image = torch.cat(image, 0).t().reshape(C_, H_, W_)
    point_weight = torch.zeros((N,), dtype=dt, device=dev)
    point_weight[order] = pweight_sorted
    D = int(((pr['rect'][:, 2] - pr['rect'][:, 0]) * (pr['rect'][:, 3] - pr['rect'][:, 1])).sum())
    return dict(image=image, radii=pr['radius'].to(torch.int32), point_id_pixel=pid.reshape(H_, W_),
                point_weight_pixel=pwp.reshape(H_, W_), point_weight=point_weight,
                final_T=torch.cat(finalT).reshape(H_, W_), n_instances=D, proj=pr, colors=colors)
