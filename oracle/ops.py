"""Operator-level oracle (TEST INFRASTRUCTURE — see oracle/__init__.py).

Each function restates one reference operator of the hot path with plain
PyTorch fp32 ops and cites the reference lines it follows (paths relative to
/root/reference).  Everything is differentiable through autograd, which is how
the oracle produces reference gradients.

`align_corners`: the reference calls `grid_sample` with its default; under the
installed torch that is `align_corners=False`.  `True` reproduces
the torch-1.2 behaviour the reference was written against.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

MAX_DISP = 4                      # pwc_tf.py:16 (md=4) / :97 (d=4)
SSIM_C1 = 0.01 ** 2               # ssim.py:5
SSIM_C2 = 0.03 ** 2               # ssim.py:6
MASK_THRESHOLD = 0.9999           # net_utils.py:50
WEIGHT_SIGMA2 = 0.03              # model_flow_paper.py:126
EPS = 1e-12                       # model_flow_paper.py:97,145,192


# --------------------------------------------------------------------------- a1
def cost_volume(f1: torch.Tensor, f2: torch.Tensor, d: int = MAX_DISP) -> torch.Tensor:
    """81-displacement cost volume, channel mean.  pwc_tf.py:97-106.

    out[b, (2d+1)*i + j, y, x] = mean_c f1[b,c,y,x] * f2[b,c,y+i-d,x+j-d], zeros outside.
    """
    assert f1.shape == f2.shape                                    # pwc_tf.py:99
    B, C, H, W = f1.shape
    n = 2 * d + 1
    f2p = F.pad(f2, (d, d, d, d), value=0.0)                       # pwc_tf.py:101
    planes = []
    for k in range(n * n):
        i, j = divmod(k, n)                                        # i = dy (major), j = dx (minor): pwc_tf.py:103-104
        planes.append((f1 * f2p[:, :, i:i + H, j:j + W]).mean(1, keepdim=True))   # pwc_tf.py:105
    return torch.cat(planes, 1)                                    # pwc_tf.py:106


# ------------------------------------------------------------------------ a2/a3
def _pixel_grid(B, H, W, device):
    xs = torch.arange(W, device=device, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
    ys = torch.arange(H, device=device, dtype=torch.float32).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat((xs, ys), 1)                                  # net_utils.py:29-33


def warp_flow(x, flow, use_mask=False, align_corners=False):
    """Bilinear backward warp, zeros padding, optional validity mask.  net_utils.py:16-54."""
    B, C, H, W = x.shape
    if tuple(flow.shape) != (B, 2, H, W):                          # net_utils.py:35-36
        raise ValueError('the shape of grid {0} is not equal to the shape of flow {1}.'.format(
            torch.Size((B, 2, H, W)), flow.shape))
    v = _pixel_grid(B, H, W, x.device) + flow                      # net_utils.py:39
    gx = 2.0 * v[:, 0] / max(W - 1, 1) - 1.0                       # net_utils.py:42
    gy = 2.0 * v[:, 1] / max(H - 1, 1) - 1.0                       # net_utils.py:43
    grid = torch.stack((gx, gy), -1)                               # net_utils.py:45
    out = F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros',
                        align_corners=align_corners)               # net_utils.py:46
    if not use_mask:
        return out
    cover = F.grid_sample(torch.ones_like(x), grid, mode='bilinear', padding_mode='zeros',
                          align_corners=align_corners)             # net_utils.py:48-49
    cover = (cover.detach() >= MASK_THRESHOLD).to(x.dtype)         # net_utils.py:50-51
    return out * cover                                             # net_utils.py:52


def sample_coords(flow, align_corners=False):
    """Source pixel coordinates (ix, iy) that `warp_flow` samples, with the same fp32
    operation order as net_utils.py:39-43 followed by ATen's grid un-normalisation."""
    B, _, H, W = flow.shape
    v = _pixel_grid(B, H, W, flow.device) + flow
    gx = 2.0 * v[:, 0] / max(W - 1, 1) - 1.0
    gy = 2.0 * v[:, 1] / max(H - 1, 1) - 1.0
    if align_corners:
        ix = (gx + 1.0) / 2.0 * (W - 1)
        iy = (gy + 1.0) / 2.0 * (H - 1)
    else:
        ix = ((gx + 1.0) * W - 1.0) / 2.0
        iy = ((gy + 1.0) * H - 1.0) / 2.0
    return ix, iy


def warp_closed_form(x, flow, use_mask=False, align_corners=False):
    """Independent gather-based statement of the same warp.  Used by the
    tests to cross-check `warp_flow` and to expose corner indices and the coverage sum."""
    B, C, H, W = x.shape
    ix, iy = sample_coords(flow, align_corners)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    tx = ix - x0
    ty = iy - y0
    x0 = x0.long()
    y0 = y0.long()
    out = torch.zeros_like(x)
    cover = torch.zeros(B, H, W, dtype=x.dtype, device=x.device)
    flat = x.reshape(B, C, H * W)
    for dy, dx, w in ((0, 0, (1 - tx) * (1 - ty)), (0, 1, tx * (1 - ty)),
                      (1, 0, (1 - tx) * ty), (1, 1, tx * ty)):
        xc = x0 + dx
        yc = y0 + dy
        inb = ((xc >= 0) & (xc < W) & (yc >= 0) & (yc < H)).to(x.dtype)
        idx = (yc.clamp(0, H - 1) * W + xc.clamp(0, W - 1)).view(B, 1, H * W).expand(B, C, H * W)
        vals = torch.gather(flat, 2, idx).view(B, C, H, W)
        out = out + vals * (w * inb).unsqueeze(1)
        cover = cover + w * inb
    if use_mask:
        out = out * (cover.detach() >= MASK_THRESHOLD).to(x.dtype).unsqueeze(1)
    return out, cover


# --------------------------------------------------------------------------- a6
def _box3(t):
    return F.avg_pool2d(t, 3, 1, padding=1)                        # ssim.py:8 (zero pad, /9 always)


def ssim(x, y):
    """SSIM map with 3x3 box moments.  pytorch_ssim/ssim.py:4-19."""
    mu_x, mu_y = _box3(x), _box3(y)
    sig_x = _box3(x * x) - mu_x * mu_x                             # ssim.py:11
    sig_y = _box3(y * y) - mu_y * mu_y                             # ssim.py:12
    sig_xy = _box3(x * y) - mu_x * mu_y                            # ssim.py:13
    num = (2 * mu_x * mu_y + SSIM_C1) * (2 * sig_xy + SSIM_C2)     # ssim.py:15
    den = (mu_x * mu_x + mu_y * mu_y + SSIM_C1) * (sig_x + sig_y + SSIM_C2)   # ssim.py:16
    return num / den


# --------------------------------------------------------------------------- a9
def img_pyramid(img, n):
    """model_flow_paper.py:54-60 (no gradient: `.data`)."""
    H, W = img.shape[2], img.shape[3]
    return [F.adaptive_avg_pool2d(img, [int(H / 2 ** s), int(W / 2 ** s)]).detach() for s in range(n)]


def warp_pyramid(imgs, flows, align_corners=False):
    """model_flow_paper.py:62-66."""
    return [warp_flow(i, f, use_mask=True, align_corners=align_corners) for i, f in zip(imgs, flows)]


# --------------------------------------------------------------------------- a4
def diff_weight(warped_l, imgs, warped_r, num_scales=3):
    """Photometric difference + soft weight ("occlusion") maps.  model_flow_paper.py:101-134.
    Returns (diff_bwd, diff_fwd, weight_bwd, weight_fwd): bwd <-> left image, fwd <-> right."""
    d_b, d_f, w_b, w_f = [], [], [], []
    for s in range(num_scales):
        wl, im, wr = warped_l[s], imgs[s], warped_r[s]
        valid_f = 1 - (wr == 0).all(1, keepdim=True).to(wr.dtype)  # :111
        valid_b = 1 - (wl == 0).all(1, keepdim=True).to(wl.dtype)  # :112
        dl = (im - wl).abs().mean(1, keepdim=True)                 # :117
        dr = (im - wr).abs().mean(1, keepdim=True)                 # :118
        a = 1 - torch.softmax(torch.cat((dl, dr), 1), 1)           # :120-121
        a = a.detach()                                             # :122
        a = 2 * torch.exp(-(a - 0.5) ** 2 / WEIGHT_SIGMA2)         # :126
        w_b.append(a[:, 0:1] * valid_b)                            # :128
        w_f.append(a[:, 1:2] * valid_f)                            # :129
        d_f.append(dr)
        d_b.append(dl)
    return d_b, d_f, w_b, w_f


# --------------------------------------------------------------------------- a5
def loss_with_mask(diffs, weights, num_scales=3):
    """model_flow_paper.py:90-99.  `diff` has one channel; the reference's repeat(1,3,1,1)
    of the mask then broadcasts diff to three identical channels before the mean."""
    total = 0
    for s in range(num_scales):
        d, w = diffs[s], weights[s]
        divider = w.mean((1, 2, 3))
        total = total + (d * w.repeat(1, 3, 1, 1)).mean((1, 2, 3)) / (divider + EPS)
    return total


def loss_ssim(imgs, warped, weights, num_scales=3):
    """model_flow_paper.py:137-148."""
    total = 0
    for s in range(num_scales):
        im, wp, w = imgs[s], warped[s], weights[s]
        divider = w.mean((1, 2, 3))
        w3 = w.repeat(1, 3, 1, 1)
        s_map = ssim(im * w3, wp * w3)
        total = total + torch.clamp((1.0 - s_map) / 2.0, 0, 1).mean((1, 2, 3)) / (divider + EPS)
    return total


# --------------------------------------------------------------------------- a7
def _grad2_error(flow, img):
    """model_flow_paper.py:152-166."""
    gx = img[:, :, :, 1:] - img[:, :, :, :-1]
    gy = img[:, :, 1:, :] - img[:, :, :-1, :]
    wx = torch.exp(-10.0 * gx.abs().mean(1, keepdim=True))
    wy = torch.exp(-10.0 * gy.abs().mean(1, keepdim=True))
    # second differences in the reference's rounding order: (f[x+2]-f[x+1]) - (f[x+1]-f[x])
    fxx = (flow[:, :, :, 2:] - flow[:, :, :, 1:-1]) - (flow[:, :, :, 1:-1] - flow[:, :, :, :-2])
    fyy = (flow[:, :, 2:, :] - flow[:, :, 1:-1, :]) - (flow[:, :, 1:-1, :] - flow[:, :, :-2, :])
    return ((wx[:, :, :, 1:] * fxx.abs()).mean((1, 2, 3)) + (wy[:, :, 1:, :] * fyy.abs()).mean((1, 2, 3))) / 2.0


def loss_flow_smooth(flows, imgs, num_scales=3):
    """model_flow_paper.py:168-177 (flow/20)."""
    total = 0
    for s in range(num_scales):
        total = total + _grad2_error(flows[s] / 20.0, imgs[s])
    return total


# --------------------------------------------------------------------------- a8
def flow_normalization(flow):
    """model_flow_paper.py:44-51."""
    return flow / (torch.norm(flow, p=2, dim=1, keepdim=True) + EPS)


def loss_flow_consis(flows_fwd, flows_bwd, weights_fwd, num_scales=3):
    """model_flow_paper.py:180-195 (bwd detached, occ = 1 - weight)."""
    total = 0
    for s in range(num_scales):
        nf = flow_normalization(flows_fwd[s])
        nb = flow_normalization(flows_bwd[s]).detach()
        occ = 1 - weights_fwd[s]
        divider = occ.mean((1, 2, 3))
        total = total + ((nf + nb).abs() * occ).mean((1, 2, 3)) / (divider + EPS)
    return total


def total_loss(loss_pack, w_ssim=0.85, w_smooth=10.0, w_consis=0.01):
    """train.py:147-150 with config_utils.py:3-9 and kitti.yaml:17-19."""
    return ((1 - w_ssim) * loss_pack['loss_pixel'].mean() + w_ssim * loss_pack['loss_ssim'].mean()
            + w_smooth * loss_pack['loss_flow_smooth'].mean() + w_consis * loss_pack['loss_flow_consis'].mean())


# ----------------------------------------------------------------------- a12/a13
# NOT IN THE REFERENCE -> parity unpinned.  Restates the published
# forward-splat of Wang et al. (occlusion-aware unsupervised flow) as four scatter_adds.
def splat_targets(flow_nhwc):
    """Integer contract of the splat: for every source pixel the four flat target indices
    b*H*W + yc*W + xc (corner order (x0,y0),(x0,y1),(x1,y0),(x1,y1)), their in-bounds flags
    and bilinear weights.  tx = x + fx, ty = y + fy in plain fp32."""
    B, H, W, _ = flow_nhwc.shape
    dev = flow_nhwc.device
    xs = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W)
    ys = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1)
    tx = xs + flow_nhwc[..., 0]
    ty = ys + flow_nhwc[..., 1]
    x0f, y0f = torch.floor(tx), torch.floor(ty)
    x1f, y1f = x0f + 1, y0f + 1
    wa = (x1f - tx) * (y1f - ty)
    wb = (x1f - tx) * (ty - y0f)
    wc = (tx - x0f) * (y1f - ty)
    wd = (tx - x0f) * (ty - y0f)
    x0, y0 = x0f.long(), y0f.long()
    base = (torch.arange(B, device=dev) * H * W).view(B, 1, 1)
    idx, inb = [], []
    for xc, yc in ((x0, y0), (x0, y0 + 1), (x0 + 1, y0), (x0 + 1, y0 + 1)):
        ok = (xc >= 0) & (xc < W) & (yc >= 0) & (yc < H)
        idx.append(torch.where(ok, base + yc * W + xc, torch.full_like(xc, -1)))
        inb.append(ok)
    return torch.stack(idx, -1), torch.stack(inb, -1), torch.stack((wa, wb, wc, wd), -1)


def splat(u_nhwc, flow_nhwc):
    """transformerFwd: out[b,yc,xc,:] += U[b,y,x,:] * w for the four corners.  (App. D)"""
    B, H, W, C = u_nhwc.shape
    idx, inb, wts = splat_targets(flow_nhwc)
    out = torch.zeros(B * H * W, C, dtype=u_nhwc.dtype, device=u_nhwc.device)
    src = u_nhwc.reshape(B * H * W, C)
    for k in range(4):
        w = (wts[..., k] * inb[..., k].to(wts.dtype)).reshape(-1, 1)
        tgt = idx[..., k].clamp_min(0).reshape(-1, 1).expand(-1, C)
        out = out.scatter_add(0, tgt, src * w)
    return out.view(B, H, W, C)


def range_map(flow_nhwc):
    return splat(torch.ones_like(flow_nhwc[..., :1]), flow_nhwc)


def occlusion_mask(flow_nhwc):
    """Soft visibility clamp(R, 0, 1) (no grad upstream)."""
    return range_map(flow_nhwc).detach().clamp(0, 1)


def fb_consistency_mask(flow_fwd, flow_rev, alpha=3.0, beta=0.05, align_corners=False):
    """noc = |f_fwd + warp(f_rev, f_fwd)|_2 < max(alpha, beta*|f_fwd|_2); NCHW flows.  (App. D,
    alpha/beta: config/kitti.yaml:26-27, read at model_flow_paper.py:29-30, unused there)."""
    back = warp_flow(flow_rev, flow_fwd, use_mask=False, align_corners=align_corners)
    diff = torch.norm(back + flow_fwd, p=2, dim=1, keepdim=True)
    bound = torch.clamp(beta * torch.norm(flow_fwd, p=2, dim=1, keepdim=True), min=alpha)
    return (diff < bound).to(flow_fwd.dtype)
