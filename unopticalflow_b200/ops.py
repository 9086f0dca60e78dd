"""Host-side operators: the reference's operator API backed by libuof_b200.so.

Names, argument meaning and error behaviour mirror the reference seams:
`corr(f1, f2)` == PWC_tf.corr_naive (pwc_tf.py:97-106), `warp_flow(x, flow, use_mask)` ==
net_utils.warp_flow (net_utils.py:16-54), `SSIM(x, y)` == pytorch_ssim.SSIM (ssim.py:4-19); the
loss functions take the same lists-of-tensors as Model_flow.compute_* (model_flow_paper.py:90-195).
Every op is a `torch.autograd.Function` whose forward/backward enqueue hand-written sm_100a kernels
on the current CUDA stream.  CPU tensors raise: there is no fallback path.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import ConsisLevel, PhotoLevel, PhotoWarpLevel, SmoothLevel

# grid_sample convention of "the reference executed under the installed torch" (SURVEY F4).
DEFAULT_ALIGN_CORNERS = False
# fp32 rounding of the sampling-coordinate chain (include/uof_b200.h, UOF_COORD_HOST): 'cuda' = what ATen's CUDA kernels
# compute (the reference as train.py runs it: reciprocal multiply + FMA), 'host' = what ATen's CPU kernels compute (true
# division; the arithmetic of the CPU oracle and the golden fixtures).  One ulp of the normalised coordinate apart, which at
# W ~ 800 is 2.5e-5 px and ~1e-4 relative in the flow gradient -- as far as the reference is from itself across devices.
COORD_ARITHMETIC = 'cuda'
# Model_flow.forward evaluates the image warps inside the photometric kernels (uof_photo_warp_loss_*); UOF_NO_PHOTO_WARP=1
# keeps the separate uof_warp_* + uof_photo_loss_* launches (A/B measurements, profiles/r2_photo_warp.md).
FUSE_IMAGE_WARP = os.environ.get('UOF_NO_PHOTO_WARP') is None


def _coord_flags(align_corners):
    if COORD_ARITHMETIC not in ('cuda', 'host'):
        raise ValueError("ops.COORD_ARITHMETIC must be 'cuda' or 'host'")
    return (1 if align_corners else 0) | (2 if COORD_ARITHMETIC == 'host' else 0)
NUM_DISPLACEMENTS = 81


def _require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('unopticalflow_b200 operators need CUDA tensors (got device %s); '
                               'there is no CPU fallback' % t.device)
        if t.dtype != torch.float32:
            raise TypeError('unopticalflow_b200 operators are fp32 only (got %s)' % t.dtype)


def _alert_not_deterministic(what):
    """Mirror of ATen's `globalContext().alertNotDeterministic`: the kernels named here combine partial results with fp32
    atomics (per-block `RED`s of the fused loss sums, the scattered d/dx of the warp, the splat, split-K > 8 of the cost
    volume), so their last bits depend on the order the blocks retire.  Under `torch.use_deterministic_algorithms(True)`
    that is an error (a warning with `warn_only=True`), exactly like `grid_sampler_2d_backward_cuda` in the reference path."""
    if torch.are_deterministic_algorithms_enabled():
        msg = ('%s does not have a deterministic implementation (fp32 atomics), but you set '
               "'torch.use_deterministic_algorithms(True)'." % what)
        if torch.is_deterministic_algorithms_warn_only_enabled():
            import warnings
            warnings.warn(msg)
        else:
            raise RuntimeError(msg)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# ------------------------------------------------------------------------------------------ a1
class _CostVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2):
        f1c, f2c = f1.contiguous(), f2.contiguous()
        B, C, H, W = f1c.shape
        out = torch.empty((B, NUM_DISPLACEMENTS, H, W), device=f1.device, dtype=torch.float32)
        with torch.cuda.device_of(f1c):
            _lib.call('uof_cost_volume_fwd', _p(f1c), _p(f2c), _p(out), B, C, H, W, NUM_DISPLACEMENTS * H * W, _stream(f1c))
        ctx.save_for_backward(f1c, f2c)
        return out

    @staticmethod
    def backward(ctx, gout):
        f1c, f2c = ctx.saved_tensors
        B, C, H, W = f1c.shape
        gout = gout.contiguous()
        g1, g2 = torch.empty_like(f1c), torch.empty_like(f2c)
        with torch.cuda.device_of(f1c):
            _lib.call('uof_cost_volume_bwd', _p(gout), NUM_DISPLACEMENTS * H * W, _p(f1c), _p(f2c), _p(g1), _p(g2),
                      B, C, H, W, _stream(f1c))
        return g1, g2


def corr(input1: torch.Tensor, input2: torch.Tensor) -> torch.Tensor:
    """81-displacement cost volume, channel mean.  Drop-in for `PWC_tf.corr` (pwc_tf.py:19,97-106)."""
    assert input1.shape == input2.shape                       # pwc_tf.py:99
    _require_cuda(input1, input2)
    return _CostVolume.apply(input1, input2)


class _CostVolumeConcat(torch.autograd.Function):
    """cat((corr(c1, c2w), c1, up_flow), 1) with the cost volume written straight into the concat buffer (forward) and
    its gradient read in place from the buffer's gradient (backward): SURVEY 8(f) "decoder glue fusion".  Saves the copy of
    the (B,81,h,w) cost volume into the concat on the way forward and the .contiguous() of its gradient slice on the way
    back (pwc_tf.py:122-123 and the matching lines of every decoder level)."""

    @staticmethod
    def forward(ctx, c1, c2w, up_flow):
        c1c, c2c = c1.contiguous(), c2w.contiguous()
        B, C, H, W = c1c.shape
        ctot = NUM_DISPLACEMENTS + C + 2
        x = torch.empty((B, ctot, H, W), device=c1.device, dtype=torch.float32)
        with torch.cuda.device_of(c1c):
            _lib.call('uof_cost_volume_fwd', _p(c1c), _p(c2c), _p(x), B, C, H, W, ctot * H * W, _stream(c1c))
        x[:, NUM_DISPLACEMENTS:NUM_DISPLACEMENTS + C].copy_(c1c)
        x[:, NUM_DISPLACEMENTS + C:].copy_(up_flow)
        ctx.save_for_backward(c1c, c2c)
        return x

    @staticmethod
    def backward(ctx, gx):
        c1c, c2c = ctx.saved_tensors
        B, C, H, W = c1c.shape
        ctot = NUM_DISPLACEMENTS + C + 2
        gx = gx.contiguous()
        g1, g2 = torch.empty_like(c1c), torch.empty_like(c2c)
        with torch.cuda.device_of(c1c):
            _lib.call('uof_cost_volume_bwd', _p(gx), ctot * H * W, _p(c1c), _p(c2c), _p(g1), _p(g2), B, C, H, W, _stream(c1c))
        g1 += gx[:, NUM_DISPLACEMENTS:NUM_DISPLACEMENTS + C]
        return g1, g2, gx[:, NUM_DISPLACEMENTS + C:]


def corr_concat(c1: torch.Tensor, c2_warped: torch.Tensor, up_flow: torch.Tensor) -> torch.Tensor:
    """== torch.cat((corr(c1, c2_warped), c1, up_flow), 1) (pwc_tf.py:122-123), fused."""
    assert c1.shape == c2_warped.shape and up_flow.shape[1] == 2 and up_flow.shape[2:] == c1.shape[2:]
    _require_cuda(c1, c2_warped, up_flow)
    return _CostVolumeConcat.apply(c1, c2_warped, up_flow)


class _DecoderInput(torch.autograd.Function):
    """Input of one decoder level, pwc_tf.py:119-123 (and :132-136, :144-148, :157-161), as ONE autograd node:

        up = 2 * F.interpolate(flow_prev, scale_factor=2, mode='bilinear')
        x  = torch.cat((corr(c1, warp(c2, up)), c1, up), 1)            -> (x, up)

    SURVEY 8(f) rank 2, "decoder glue fusion".  `c` may carry 1/rep of the batch of `c2` (the training step's first
    feature list is [centre; centre]): it is stored ONCE, broadcast into the `c1` slice of the concat buffer, and the cost
    volume reads its first operand in place from that slice (batch stride of the buffer) -- no torch.cat((c, c)) and no
    copy of c1 into the concat.  The up-sampling kernel writes `up` both densely (for the warp and the residual add) and
    into the buffer's last two channels.  Backward: the cost-volume kernel adds the `c1` slice gradient in its epilogue,
    the up-sampling backward sums its three gradient sources (warp, concat slice, residual) while it gathers."""

    @staticmethod
    def forward(ctx, c, c2, flow_prev, align_corners):
        c, c2, fp = c.contiguous(), c2.contiguous(), flow_prev.contiguous()
        B, C, H, W = c2.shape
        B1 = c.shape[0]
        assert B % B1 == 0 and tuple(c.shape[1:]) == (C, H, W) and tuple(fp.shape) == (B, 2, H // 2, W // 2)
        nd = NUM_DISPLACEMENTS
        ctot, plane = nd + C + 2, H * W
        dev = c2.device
        x = torch.empty((B, ctot, H, W), device=dev, dtype=torch.float32)
        up = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
        c2w = torch.empty_like(c2)
        flags = _coord_flags(align_corners)
        elt = x.element_size()
        with torch.cuda.device_of(c2):
            st = _stream(c2)
            _lib.call('uof_upsample_bilinear_fwd2', _p(fp), _p(up), ctypes.c_void_p(x.data_ptr() + (nd + C) * plane * elt), 2,
                      ctot * plane, B * 2, H // 2, W // 2, H, W, 2.0, st)
            _lib.call('uof_warp_fwd', _p(c2), _p(up), _p(c2w), B, C, H, W, 0, flags, 0, st)
            # c -> the c1 slice of every replica (one broadcasting copy)
            x.view(B // B1, B1, ctot, H, W)[:, :, nd:nd + C].copy_(c.unsqueeze(0).expand(B // B1, B1, C, H, W))
            _lib.call('uof_cost_volume_fwd_ex', ctypes.c_void_p(x.data_ptr() + nd * plane * elt), ctot * plane, _p(c2w), _p(x),
                      B, C, H, W, ctot * plane, st)
        ctx.save_for_backward(x, c2, c2w, up)
        ctx.dims = (B, B1, C, H, W, flags)
        return x, up

    @staticmethod
    def backward(ctx, gx, g_up):
        x, c2, c2w, up = ctx.saved_tensors
        B, B1, C, H, W, flags = ctx.dims
        nd = NUM_DISPLACEMENTS
        ctot, plane = nd + C + 2, H * W
        dev = x.device
        if gx is None:
            gx = torch.zeros_like(x)
        gx = gx.contiguous()
        if g_up is not None:
            g_up = g_up.contiguous()
        elt = x.element_size()
        g1, g2w = torch.empty((B, C, H, W), device=dev, dtype=torch.float32), torch.empty_like(c2w)
        gc2, gflow = torch.empty_like(c2), torch.empty_like(up)
        gprev = torch.empty((B, 2, H // 2, W // 2), device=dev, dtype=torch.float32)
        _alert_not_deterministic('uof_warp_bwd (gradient w.r.t. the warped tensor)')
        with torch.cuda.device_of(x):
            st = _stream(x)
            _lib.call('uof_cost_volume_bwd_ex', _p(gx), ctot * plane, ctypes.c_void_p(x.data_ptr() + nd * plane * elt), ctot * plane,
                      _p(c2w), ctypes.c_void_p(gx.data_ptr() + nd * plane * elt), ctot * plane, _p(g1), _p(g2w), B, C, H, W, st)
            _lib.call('uof_warp_bwd', _p(g2w), _p(c2), _p(up), _p(gc2), _p(gflow), B, C, H, W, 0, flags, 0, st)
            _lib.call('uof_upsample_bilinear_bwd3', _p(gflow), ctypes.c_void_p(gx.data_ptr() + (nd + C) * plane * elt), 2,
                      ctot * plane, _p(g_up), _p(gprev), B * 2, H // 2, W // 2, H, W, 2.0, st)
        if B1 == B:
            gc = g1
        elif B == 2 * B1:           # the training step: one vectorised add of the two halves (ATen's reduce kernel is 2x slower here)
            gc = g1[:B1] + g1[B1:]
        else:
            gc = g1.view(B // B1, B1, C, H, W).sum(0)
        return gc, gc2, gprev, None


def decoder_input(c: torch.Tensor, c2: torch.Tensor, flow_prev: torch.Tensor, align_corners: bool | None = None):
    """-> (x, up) with up = 2 * interpolate(flow_prev, x2) and x = cat((corr(c1, warp(c2, up)), c1, up), 1), where c1 is
    `c` repeated along the batch to c2's batch size (pwc_tf.py:119-123), fused (see _DecoderInput)."""
    _require_cuda(c, c2, flow_prev)
    B, C, H, W = c2.shape
    if (c2.shape[0] % c.shape[0] or tuple(c.shape[1:]) != (C, H, W)
            or tuple(flow_prev.shape) != (B, 2, H // 2, W // 2) or H % 2 or W % 2):
        raise ValueError('decoder_input: c %r, c2 %r and flow_prev %r do not form a decoder level'
                         % (tuple(c.shape), tuple(c2.shape), tuple(flow_prev.shape)))
    ac = DEFAULT_ALIGN_CORNERS if align_corners is None else bool(align_corners)
    return _DecoderInput.apply(c, c2, flow_prev, ac)


# --------------------------------------------------------------------------------------- a2/a3
def _is_channels_last(x):
    return (x.dim() == 4 and x.shape[1] % 4 == 0 and x.shape[1] > 1 and not x.is_contiguous()
            and x.is_contiguous(memory_format=torch.channels_last))


class _WarpFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow, use_mask, align_corners):
        cl = _is_channels_last(x)
        xc = x if cl else x.contiguous()
        fc = flow.contiguous()
        B, C, H, W = xc.shape
        out = torch.empty_like(xc)     # preserves the memory format
        with torch.cuda.device_of(xc):
            _lib.call('uof_warp_fwd', _p(xc), _p(fc), _p(out), B, C, H, W, int(use_mask), _coord_flags(align_corners), int(cl),
                      _stream(xc))
        ctx.save_for_backward(xc, fc)
        ctx.flags = (int(use_mask), _coord_flags(align_corners), int(cl))
        return out

    @staticmethod
    def backward(ctx, gout):
        xc, fc = ctx.saved_tensors
        use_mask, align_corners, cl = ctx.flags
        B, C, H, W = xc.shape
        gout = gout.contiguous(memory_format=torch.channels_last) if cl else gout.contiguous()
        gx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        if gx is not None:
            _alert_not_deterministic('uof_warp_bwd (gradient w.r.t. the warped tensor)')
        gflow = torch.empty_like(fc)
        with torch.cuda.device_of(xc):
            _lib.call('uof_warp_bwd', _p(gout), _p(xc), _p(fc), _p(gx), _p(gflow), B, C, H, W, use_mask, align_corners,
                      cl, _stream(xc))
        return gx, gflow, None, None


def warp_flow(x: torch.Tensor, flow: torch.Tensor, use_mask: bool = False, align_corners: bool | None = None):
    """Bilinear backward warp (+ validity mask).  Drop-in for `warp_flow` (net_utils.py:16-54)."""
    B, C, H, W = x.size()
    if tuple(flow.shape) != (B, 2, H, W):                     # net_utils.py:35-36
        raise ValueError('the shape of grid {0} is not equal to the shape of flow {1}.'.format(
            torch.Size((B, 2, H, W)), flow.shape))
    _require_cuda(x, flow)
    ac = DEFAULT_ALIGN_CORNERS if align_corners is None else bool(align_corners)
    return _WarpFlow.apply(x, flow, bool(use_mask), ac)


# ------------------------------------------------------------------------------------------ a6
class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        xc, yc = x.contiguous(), y.contiguous()
        B, C, H, W = xc.shape
        out = torch.empty_like(xc)
        with torch.cuda.device_of(xc):
            _lib.call('uof_ssim_fwd', _p(xc), _p(yc), _p(out), B * C, H, W, _stream(xc))
        ctx.save_for_backward(xc, yc)
        return out

    @staticmethod
    def backward(ctx, gout):
        xc, yc = ctx.saved_tensors
        B, C, H, W = xc.shape
        gx, gy = torch.empty_like(xc), torch.empty_like(yc)
        with torch.cuda.device_of(xc):
            _lib.call('uof_ssim_bwd', _p(gout.contiguous()), _p(xc), _p(yc), _p(gx), _p(gy), B * C, H, W, _stream(xc))
        return gx, gy


def SSIM(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """SSIM map.  Drop-in for `pytorch_ssim.SSIM` (ssim.py:4-19)."""
    assert x.shape == y.shape and x.dim() == 4
    _require_cuda(x, y)
    return _SSIM.apply(x, y)


# ------------------------------------------------------------------------------------ a4+a5+a6
def _levels(cls, n):
    return (cls * n)()


class _PhotoLoss(torch.autograd.Function):
    """Fused compute_diff_weight + 2x compute_loss_with_mask + 2x compute_loss_ssim.

    tensors = imgs[S] + warped_l[S] + warped_r[S], or with `stacked` imgs[S] + warped_lr[S] where
    warped_lr[s] is (2B,3,H,W) = [from-left ; from-right] (what the 2B-batched decoder produces), in
    which case the gradient comes back as one (2B,3,H,W) tensor per level."""

    @staticmethod
    def forward(ctx, S, want_diff, stacked, *tensors):
        imgs = [t.contiguous() for t in tensors[:S]]
        B = imgs[0].shape[0]
        if stacked:
            both = [t.contiguous() for t in tensors[S:2 * S]]
            wl, wr = [t[:B] for t in both], [t[B:] for t in both]
        else:
            both = []
            wl = [t.contiguous() for t in tensors[S:2 * S]]
            wr = [t.contiguous() for t in tensors[2 * S:3 * S]]
        dev = imgs[0].device
        lv = _levels(PhotoLevel, S)
        weights_l, weights_r, diffs_l, diffs_r = [], [], [], []
        for s in range(S):
            _, _, H, W = imgs[s].shape
            assert wl[s].shape == imgs[s].shape and wr[s].shape == imgs[s].shape and imgs[s].shape[1] == 3
            weights_l.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
            weights_r.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
            if want_diff:
                diffs_l.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
                diffs_r.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
            lv[s] = PhotoLevel(imgs[s].data_ptr(), wl[s].data_ptr(), wr[s].data_ptr(),
                               weights_l[s].data_ptr(), weights_r[s].data_ptr(),
                               diffs_l[s].data_ptr() if want_diff else None, diffs_r[s].data_ptr() if want_diff else None,
                               None, None, H, W)
        _alert_not_deterministic('uof_photo_loss_fwd')
        sums = torch.empty(S * B * 6 + _lib.SUMS_EXTRA, device=dev, dtype=torch.float32)      # (S,B,6) partial sums + the kernel's block counter
        loss_pixel, loss_ssim = torch.empty(B, device=dev, dtype=torch.float32), torch.empty(B, device=dev, dtype=torch.float32)
        with torch.cuda.device_of(imgs[0]):
            _lib.call('uof_photo_loss_fwd', lv, S, B, _p(sums), _p(loss_pixel), _p(loss_ssim), _stream(imgs[0]))
        # the weight maps are saved too: with them the backward kernel skips the weight math (split variant)
        if stacked:
            ctx.save_for_backward(sums, *imgs, *both, *weights_l, *weights_r)
        else:
            ctx.save_for_backward(sums, *imgs, *wl, *wr, *weights_l, *weights_r)
        ctx.S, ctx.stacked = S, stacked
        outs = (loss_pixel, loss_ssim, *weights_l, *weights_r, *diffs_l, *diffs_r)
        ctx.mark_non_differentiable(*outs[2:])
        return outs

    @staticmethod
    def backward(ctx, g_pixel, g_ssim, *unused):
        S, stacked = ctx.S, ctx.stacked
        sums, *rest = ctx.saved_tensors
        imgs = rest[:S]
        B = imgs[0].shape[0]
        if stacked:
            both = rest[S:2 * S]
            wl, wr = [t[:B] for t in both], [t[B:] for t in both]
            gboth = [torch.empty_like(t) for t in both]
            gl, gr = [t[:B] for t in gboth], [t[B:] for t in gboth]
            wmaps = rest[2 * S:4 * S]
        else:
            wl, wr = rest[S:2 * S], rest[2 * S:3 * S]
            gl, gr = [torch.empty_like(t) for t in wl], [torch.empty_like(t) for t in wr]
            wmaps = rest[3 * S:5 * S]
        lv = _levels(PhotoLevel, S)
        for s in range(S):
            _, _, H, W = imgs[s].shape
            lv[s] = PhotoLevel(imgs[s].data_ptr(), wl[s].data_ptr(), wr[s].data_ptr(), wmaps[s].data_ptr(),
                               wmaps[S + s].data_ptr(), None, None, gl[s].data_ptr(), gr[s].data_ptr(), H, W)
        g_pixel = torch.zeros(B, device=sums.device, dtype=torch.float32) if g_pixel is None else g_pixel.contiguous()
        g_ssim = torch.zeros(B, device=sums.device, dtype=torch.float32) if g_ssim is None else g_ssim.contiguous()
        with torch.cuda.device_of(sums):
            _lib.call('uof_photo_loss_bwd', lv, S, B, _p(sums), _p(g_pixel), _p(g_ssim), _stream(sums))
        if stacked:
            return (None, None, None, *([None] * S), *gboth)
        return (None, None, None, *([None] * S), *gl, *gr)


def _unpack_photo(outs, S, return_diffs):
    res = (outs[0], outs[1], list(outs[2:2 + S]), list(outs[2 + S:2 + 2 * S]))
    if return_diffs:
        res += (list(outs[2 + 2 * S:2 + 3 * S]), list(outs[2 + 3 * S:2 + 4 * S]))
    return res


def photometric_losses(img_pyramid, warped_from_l, warped_from_r, num_scales=3, return_diffs=False):
    """Fused photometric path of Model_flow.forward (model_flow_paper.py:240-245).

    Returns (loss_pixel (B,), loss_ssim (B,), weight_bwd list, weight_fwd list[, diff_bwd, diff_fwd]):
    "bwd" pairs with the left image, "fwd" with the right one, as in the reference."""
    S = num_scales
    _require_cuda(*img_pyramid[:S], *warped_from_l[:S], *warped_from_r[:S])
    outs = _PhotoLoss.apply(S, bool(return_diffs), False, *img_pyramid[:S], *warped_from_l[:S], *warped_from_r[:S])
    return _unpack_photo(outs, S, return_diffs)


def photometric_losses_stacked(img_pyramid, warped_lr, num_scales=3, return_diffs=False):
    """Same, with the two warped pyramids stacked along the batch: warped_lr[s] = (2B,3,H,W)."""
    S = num_scales
    _require_cuda(*img_pyramid[:S], *warped_lr[:S])
    outs = _PhotoLoss.apply(S, bool(return_diffs), True, *img_pyramid[:S], *warped_lr[:S])
    return _unpack_photo(outs, S, return_diffs)


# --------------------------------------------------------------------------------- a3+a4+a5+a6
class _PhotoWarpLoss(torch.autograd.Function):
    """Image warps fused into the photometric losses: warp_flow_pyramid x2 + compute_diff_weight + 2x compute_loss_with_mask
    + 2x compute_loss_ssim (model_flow_paper.py:236-245) as ONE launch each way.

    tensors = imgs[S] + sources[S] + flows[S]: imgs[s] (B,3,H,W) target, sources[s] (2B,3,H,W) = [left ; right] images,
    flows[s] (2B,2,H,W) = [target->left ("bwd") ; target->right ("fwd")].  Gradient flows to the flows only (the images
    are data).  Outputs: loss_pixel, loss_ssim, weight maps (l then r), [diff maps], masked warped images (2B,3,H,W) --
    everything after the two losses is non-differentiable, as in the reference (weights are detached, :131-132)."""

    @staticmethod
    def forward(ctx, S, flags, want_diff, *tensors):
        imgs = [t.contiguous() for t in tensors[:S]]
        srcs = [t.contiguous() for t in tensors[S:2 * S]]
        flows = [t.contiguous() for t in tensors[2 * S:3 * S]]
        B = imgs[0].shape[0]
        dev = imgs[0].device
        lv = _levels(PhotoWarpLevel, S)
        warped, weights_l, weights_r, diffs_l, diffs_r = [], [], [], [], []
        for s in range(S):
            _, _, H, W = imgs[s].shape
            if (tuple(imgs[s].shape) != (B, 3, H, W) or tuple(srcs[s].shape) != (2 * B, 3, H, W)
                    or tuple(flows[s].shape) != (2 * B, 2, H, W)):
                raise ValueError('photometric_losses_warped level %d: image %r, sources %r, flows %r must be (B,3,H,W), '
                                 '(2B,3,H,W), (2B,2,H,W)' % (s, tuple(imgs[s].shape), tuple(srcs[s].shape), tuple(flows[s].shape)))
            warped.append(torch.empty((2 * B, 3, H, W), device=dev, dtype=torch.float32))
            weights_l.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
            weights_r.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
            if want_diff:
                diffs_l.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
                diffs_r.append(torch.empty((B, 1, H, W), device=dev, dtype=torch.float32))
            lv[s] = PhotoWarpLevel(imgs[s].data_ptr(), srcs[s][:B].data_ptr(), srcs[s][B:].data_ptr(),
                                   flows[s][:B].data_ptr(), flows[s][B:].data_ptr(),
                                   warped[s][:B].data_ptr(), warped[s][B:].data_ptr(),
                                   weights_l[s].data_ptr(), weights_r[s].data_ptr(),
                                   diffs_l[s].data_ptr() if want_diff else None, diffs_r[s].data_ptr() if want_diff else None,
                                   None, None, H, W)
        _alert_not_deterministic('uof_photo_warp_loss_fwd')
        sums = torch.empty(S * B * 6 + _lib.SUMS_EXTRA, device=dev, dtype=torch.float32)
        loss_pixel, loss_ssim = torch.empty(B, device=dev, dtype=torch.float32), torch.empty(B, device=dev, dtype=torch.float32)
        with torch.cuda.device_of(imgs[0]):
            _lib.call('uof_photo_warp_loss_fwd', lv, S, B, flags, _p(sums), _p(loss_pixel), _p(loss_ssim), _stream(imgs[0]))
        ctx.save_for_backward(sums, *imgs, *srcs, *flows, *warped, *weights_l, *weights_r)
        ctx.S, ctx.flags = S, flags
        outs = (loss_pixel, loss_ssim, *weights_l, *weights_r, *diffs_l, *diffs_r, *warped)
        ctx.mark_non_differentiable(*outs[2:])
        return outs

    @staticmethod
    def backward(ctx, g_pixel, g_ssim, *unused):
        S = ctx.S
        sums, *rest = ctx.saved_tensors
        imgs, srcs, flows, warped = rest[:S], rest[S:2 * S], rest[2 * S:3 * S], rest[3 * S:4 * S]
        wl, wr = rest[4 * S:5 * S], rest[5 * S:6 * S]
        B = imgs[0].shape[0]
        gflows = [torch.empty_like(f) for f in flows]
        lv = _levels(PhotoWarpLevel, S)
        for s in range(S):
            _, _, H, W = imgs[s].shape
            lv[s] = PhotoWarpLevel(imgs[s].data_ptr(), srcs[s][:B].data_ptr(), srcs[s][B:].data_ptr(),
                                   flows[s][:B].data_ptr(), flows[s][B:].data_ptr(),
                                   warped[s][:B].data_ptr(), warped[s][B:].data_ptr(), wl[s].data_ptr(), wr[s].data_ptr(),
                                   None, None, gflows[s][:B].data_ptr(), gflows[s][B:].data_ptr(), H, W)
        g_pixel = torch.zeros(B, device=sums.device, dtype=torch.float32) if g_pixel is None else g_pixel.contiguous()
        g_ssim = torch.zeros(B, device=sums.device, dtype=torch.float32) if g_ssim is None else g_ssim.contiguous()
        with torch.cuda.device_of(sums):
            _lib.call('uof_photo_warp_loss_bwd', lv, S, B, ctx.flags, _p(sums), _p(g_pixel), _p(g_ssim), _stream(sums))
        return (None, None, None, *([None] * (2 * S)), *gflows)


def photometric_losses_warped(img_pyramid, sources_lr, flows_lr, num_scales=3, align_corners=None, return_diffs=False,
                              return_warped=False):
    """`warp_flow_pyramid` (both directions, use_mask=True) + the fused photometric losses of Model_flow.forward
    (model_flow_paper.py:236-245) without materialising the warped pyramids between two launches.

    img_pyramid[s] (B,3,H,W); sources_lr[s] (2B,3,H,W) = [left ; right]; flows_lr[s] (2B,2,H,W) = [bwd ; fwd].
    Returns (loss_pixel (B,), loss_ssim (B,), weight_bwd list, weight_fwd list[, diff_bwd, diff_fwd][, warped_lr list]).
    Levels with odd W (the pair kernels need even W) take the separate warp + photometric kernels instead."""
    S = num_scales
    imgs, srcs, flows = list(img_pyramid[:S]), list(sources_lr[:S]), list(flows_lr[:S])
    _require_cuda(*imgs, *srcs, *flows)
    ac = DEFAULT_ALIGN_CORNERS if align_corners is None else bool(align_corners)
    if any(t.requires_grad for t in imgs + srcs):
        raise ValueError('photometric_losses_warped: the images are data (no gradient w.r.t. them); use warp_flow + '
                         'photometric_losses_stacked for differentiable sources')
    if any(int(t.shape[-1]) % 2 for t in imgs):
        warped = [warp_flow(srcs[s], flows[s], use_mask=True, align_corners=ac) for s in range(S)]
        res = photometric_losses_stacked(imgs, warped, S, return_diffs)
        return res + (warped,) if return_warped else res
    outs = _PhotoWarpLoss.apply(S, _coord_flags(ac), bool(return_diffs), *imgs, *srcs, *flows)
    res = _unpack_photo(outs, S, return_diffs)
    if return_warped:
        res += (list(outs[-S:]),)
    return res


class _FlowLossPack(torch.autograd.Function):
    """All four losses of Model_flow.forward (model_flow_paper.py:236-251) as ONE autograd node: the fused warp +
    photometric launch, the smoothness of both directions and the consistency loss read the same flow pyramid, and with
    three separate nodes autograd sums their three flow gradients with ATen kernels (per level: a zero-filled full-size
    tensor + a copy for the consistency loss's [B:] slice and two adds -- 12 launches, ~0.1 ms per step).  Here the
    backward pass runs the three backward kernels into ONE gradient buffer per level: uof_photo_warp_loss_bwd writes it,
    uof_smooth_loss_bwd_acc and uof_consis_loss_bwd_acc add to it in place.

    tensors = imgs[S] + sources[S] + flows[S] as in _PhotoWarpLoss.  Outputs: loss_pixel (B), loss_ssim (B),
    loss_smooth (B) = forward-flow term + backward-flow term, loss_consis (B), then the weight maps (l then r;
    non-differentiable)."""

    @staticmethod
    def forward(ctx, S, flags, *tensors):
        imgs = [t.contiguous() for t in tensors[:S]]
        srcs = [t.contiguous() for t in tensors[S:2 * S]]
        flows = [t.contiguous() for t in tensors[2 * S:3 * S]]
        B = imgs[0].shape[0]
        dev = imgs[0].device
        f32 = dict(device=dev, dtype=torch.float32)
        pl, sl, cl = _levels(PhotoWarpLevel, S), _levels(SmoothLevel, S), _levels(ConsisLevel, S)
        warped, wl, wr = [], [], []
        for s in range(S):
            _, _, H, W = imgs[s].shape
            if (tuple(imgs[s].shape) != (B, 3, H, W) or tuple(srcs[s].shape) != (2 * B, 3, H, W)
                    or tuple(flows[s].shape) != (2 * B, 2, H, W)):
                raise ValueError('flow_loss_pack level %d: image %r, sources %r, flows %r must be (B,3,H,W), (2B,3,H,W), '
                                 '(2B,2,H,W)' % (s, tuple(imgs[s].shape), tuple(srcs[s].shape), tuple(flows[s].shape)))
            warped.append(torch.empty((2 * B, 3, H, W), **f32))
            wl.append(torch.empty((B, 1, H, W), **f32))
            wr.append(torch.empty((B, 1, H, W), **f32))
            pl[s] = PhotoWarpLevel(imgs[s].data_ptr(), srcs[s][:B].data_ptr(), srcs[s][B:].data_ptr(),
                                   flows[s][:B].data_ptr(), flows[s][B:].data_ptr(), warped[s][:B].data_ptr(),
                                   warped[s][B:].data_ptr(), wl[s].data_ptr(), wr[s].data_ptr(), None, None, None, None, H, W)
            sl[s] = SmoothLevel(flows[s].data_ptr(), imgs[s].data_ptr(), None, H, W)
            cl[s] = ConsisLevel(flows[s][B:].data_ptr(), flows[s][:B].data_ptr(), wr[s].data_ptr(), None, H, W)
        _alert_not_deterministic('uof_photo_warp_loss_fwd / uof_smooth_loss_fwd / uof_consis_loss_fwd')
        psums = torch.empty(S * B * 6 + _lib.SUMS_EXTRA, **f32)
        ssums = torch.empty(S * 2 * B * 2 + _lib.SUMS_EXTRA, **f32)
        csums = torch.empty(S * B * 2 + _lib.SUMS_EXTRA, **f32)
        loss_pixel, loss_ssim, consis = torch.empty(B, **f32), torch.empty(B, **f32), torch.empty(B, **f32)
        smooth = torch.empty(2 * B, **f32)
        st = _stream(imgs[0])
        with torch.cuda.device_of(imgs[0]):
            _lib.call('uof_photo_warp_loss_fwd', pl, S, B, flags, _p(psums), _p(loss_pixel), _p(loss_ssim), st)
            _lib.call('uof_smooth_loss_fwd', sl, S, 2 * B, B, _p(ssums), _p(smooth), st)
            _lib.call('uof_consis_loss_fwd', cl, S, B, _p(csums), _p(consis), st)
        ctx.save_for_backward(psums, csums, *imgs, *srcs, *flows, *warped, *wl, *wr)
        ctx.S, ctx.flags = S, flags
        # model_flow_paper.py:248-249: the smoothness of the forward flow plus that of the backward flow
        outs = (loss_pixel, loss_ssim, smooth[B:] + smooth[:B], consis, *wl, *wr)
        ctx.mark_non_differentiable(*outs[4:])
        return outs

    @staticmethod
    def backward(ctx, g_pixel, g_ssim, g_smooth, g_consis, *unused):
        S = ctx.S
        psums, csums, *rest = ctx.saved_tensors
        imgs, srcs, flows, warped = rest[:S], rest[S:2 * S], rest[2 * S:3 * S], rest[3 * S:4 * S]
        wl, wr = rest[4 * S:5 * S], rest[5 * S:6 * S]
        B = imgs[0].shape[0]
        dev = psums.device
        gflows = [torch.empty_like(f) for f in flows]
        pl, sl, cl = _levels(PhotoWarpLevel, S), _levels(SmoothLevel, S), _levels(ConsisLevel, S)
        for s in range(S):
            _, _, H, W = imgs[s].shape
            pl[s] = PhotoWarpLevel(imgs[s].data_ptr(), srcs[s][:B].data_ptr(), srcs[s][B:].data_ptr(),
                                   flows[s][:B].data_ptr(), flows[s][B:].data_ptr(), warped[s][:B].data_ptr(),
                                   warped[s][B:].data_ptr(), wl[s].data_ptr(), wr[s].data_ptr(), None, None,
                                   gflows[s][:B].data_ptr(), gflows[s][B:].data_ptr(), H, W)
            sl[s] = SmoothLevel(flows[s].data_ptr(), imgs[s].data_ptr(), gflows[s].data_ptr(), H, W)
            cl[s] = ConsisLevel(flows[s][B:].data_ptr(), flows[s][:B].data_ptr(), wr[s].data_ptr(), gflows[s][B:].data_ptr(), H, W)
        zeros = lambda n: torch.zeros(n, device=dev, dtype=torch.float32)
        g_pixel = zeros(B) if g_pixel is None else g_pixel.contiguous()
        g_ssim = zeros(B) if g_ssim is None else g_ssim.contiguous()
        st = _stream(psums)
        with torch.cuda.device_of(psums):
            # the photometric kernel WRITES every element of the gradient buffers; the other two add to them
            _lib.call('uof_photo_warp_loss_bwd', pl, S, B, ctx.flags, _p(psums), _p(g_pixel), _p(g_ssim), st)
            if g_smooth is not None:
                _lib.call('uof_smooth_loss_bwd_acc', sl, S, 2 * B, B, _p(g_smooth.repeat(2)), 1, st)     # both directions get g
            if g_consis is not None:
                _lib.call('uof_consis_loss_bwd_acc', cl, S, B, _p(csums), _p(g_consis.contiguous()), 1, st)
        return (None, None, *([None] * (2 * S)), *gflows)


def flow_loss_pack(img_pyramid, sources_lr, flows_lr, num_scales=3, align_corners=None):
    """The loss pack of Model_flow.forward (model_flow_paper.py:236-251) from the image pyramids and the stacked flows, as
    one autograd node (see _FlowLossPack).  img_pyramid[s] (B,3,H,W); sources_lr[s] (2B,3,H,W) = [left ; right];
    flows_lr[s] (2B,2,H,W) = [bwd ; fwd].  Returns (loss_pixel (B,), loss_ssim (B,), loss_flow_smooth (B,) = forward + backward
    flow term, loss_flow_consis (B,), weight_bwd list, weight_fwd list).  Needs even W at every level (else compose
    photometric_losses_warped + flow_smooth_loss + flow_consis_loss, which is what this node computes)."""
    S = num_scales
    imgs, srcs, flows = list(img_pyramid[:S]), list(sources_lr[:S]), list(flows_lr[:S])
    _require_cuda(*imgs, *srcs, *flows)
    ac = DEFAULT_ALIGN_CORNERS if align_corners is None else bool(align_corners)
    if any(t.requires_grad for t in imgs + srcs):
        raise ValueError('flow_loss_pack: the images are data (no gradient w.r.t. them)')
    B = imgs[0].shape[0]
    if any(int(t.shape[-1]) % 2 for t in imgs):
        pix, ssim, w_b, w_f = photometric_losses_warped(imgs, srcs, flows, S, ac)
        smooth = flow_smooth_loss(flows, imgs, S)
        consis = flow_consis_loss([f[B:] for f in flows], [f[:B] for f in flows], w_f, S)
        return pix, ssim, smooth[B:] + smooth[:B], consis, w_b, w_f
    outs = _FlowLossPack.apply(S, _coord_flags(ac), *imgs, *srcs, *flows)
    return outs[0], outs[1], outs[2], outs[3], list(outs[4:4 + S]), list(outs[4 + S:4 + 2 * S])


# ------------------------------------------------------------------------------- a4 / a5 seams
class _DiffWeight(torch.autograd.Function):
    """One level of compute_diff_weight: -> diff_l, diff_r (differentiable), weight_l, weight_r (detached)."""

    @staticmethod
    def forward(ctx, img, wl, wr):
        img, wl, wr = img.contiguous(), wl.contiguous(), wr.contiguous()
        B, _, H, W = img.shape
        outs = [torch.empty((B, 1, H, W), device=img.device, dtype=torch.float32) for _ in range(4)]
        with torch.cuda.device_of(img):
            _lib.call('uof_diff_weight_fwd', _p(img), _p(wl), _p(wr), *[_p(t) for t in outs], B, H, W, _stream(img))
        ctx.save_for_backward(img, wl, wr)
        ctx.mark_non_differentiable(outs[2], outs[3])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_dl, g_dr, *unused):
        img, wl, wr = ctx.saved_tensors
        B, _, H, W = img.shape
        gl, gr = torch.empty_like(wl), torch.empty_like(wr)
        g_dl = None if g_dl is None else g_dl.contiguous()
        g_dr = None if g_dr is None else g_dr.contiguous()
        with torch.cuda.device_of(img):
            _lib.call('uof_diff_weight_bwd', _p(img), _p(wl), _p(wr), _p(g_dl), _p(g_dr), _p(gl), _p(gr), B, H, W, _stream(img))
        return None, gl, gr


def diff_weight(img_from_l, img, img_from_r, num_scales=3):
    """Drop-in for Model_flow.compute_diff_weight (model_flow_paper.py:101-134):
    -> diff_bwd, diff_fwd, weight_bwd, weight_fwd ("bwd" <-> left image, "fwd" <-> right image)."""
    d_b, d_f, w_b, w_f = [], [], [], []
    for s in range(num_scales):
        _require_cuda(img[s], img_from_l[s], img_from_r[s])
        assert img[s].shape[1] == 3 and img_from_l[s].shape == img[s].shape == img_from_r[s].shape
        dl, dr, wl, wr = _DiffWeight.apply(img[s].detach(), img_from_l[s], img_from_r[s])
        d_b.append(dl); d_f.append(dr); w_b.append(wl); w_f.append(wr)
    return d_b, d_f, w_b, w_f


class _MaskedMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, *tensors):
        diffs = [t.contiguous() for t in tensors[:S]]
        masks = [t.contiguous() for t in tensors[S:2 * S]]
        B, C = diffs[0].shape[0], diffs[0].shape[1]
        dev = diffs[0].device
        for d, m in zip(diffs, masks):
            assert d.shape[1] == C and m.shape[1] == 1 and d.shape[2:] == m.shape[2:] and d.shape[0] == m.shape[0] == B
        arr = lambda ts: (ctypes.c_void_p * S)(*[t.data_ptr() for t in ts])
        Hs = (ctypes.c_int * S)(*[d.shape[2] for d in diffs])
        Ws = (ctypes.c_int * S)(*[d.shape[3] for d in diffs])
        sums, loss = torch.empty((S, B, 2), device=dev, dtype=torch.float32), torch.empty(B, device=dev, dtype=torch.float32)
        _alert_not_deterministic('uof_masked_mean_fwd')
        with torch.cuda.device_of(diffs[0]):
            _lib.call('uof_masked_mean_fwd', arr(diffs), arr(masks), Hs, Ws, S, B, C, _p(sums), _p(loss), _stream(diffs[0]))
        ctx.save_for_backward(sums, *diffs, *masks)
        ctx.S = S
        return loss

    @staticmethod
    def backward(ctx, g):
        S = ctx.S
        sums, *rest = ctx.saved_tensors
        diffs, masks = rest[:S], rest[S:]
        B, C = diffs[0].shape[0], diffs[0].shape[1]
        gd = [torch.empty_like(d) for d in diffs]
        arr = lambda ts: (ctypes.c_void_p * S)(*[t.data_ptr() for t in ts])
        Hs = (ctypes.c_int * S)(*[d.shape[2] for d in diffs])
        Ws = (ctypes.c_int * S)(*[d.shape[3] for d in diffs])
        g = g.contiguous()
        with torch.cuda.device_of(g):
            _lib.call('uof_masked_mean_bwd', arr(diffs), arr(masks), arr(gd), Hs, Ws, S, B, C, _p(sums), _p(g), _stream(g))
        return (None, *gd, *([None] * S))


def loss_with_mask(diff_list, occ_mask_list, num_scales=3):
    """Drop-in for Model_flow.compute_loss_with_mask (model_flow_paper.py:90-99); the mask carries no gradient."""
    S = num_scales
    _require_cuda(*diff_list[:S], *occ_mask_list[:S])
    return _MaskedMean.apply(S, *diff_list[:S], *[m.detach() for m in occ_mask_list[:S]])


# ------------------------------------------------------------------------------------------ a7
class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, *tensors):
        flows = [t.contiguous() for t in tensors[:S]]
        imgs = [t.contiguous() for t in tensors[S:2 * S]]
        B, Bimg = flows[0].shape[0], imgs[0].shape[0]
        dev = flows[0].device
        lv = _levels(SmoothLevel, S)
        if Bimg <= 0 or B % Bimg:
            raise ValueError('flow_smooth_loss: the flow batch %d must be a multiple of the image batch %d' % (B, Bimg))
        for s in range(S):
            _, _, H, W = flows[s].shape
            # the kernels hard-code 2 flow and 3 image channels and index the image with the flow's sizes
            if tuple(flows[s].shape) != (B, 2, H, W) or tuple(imgs[s].shape) != (Bimg, 3, H, W):
                raise ValueError('flow_smooth_loss level %d: flow %r must be (B,2,H,W) and img %r (B_img,3,H,W) of the same size'
                                 % (s, tuple(flows[s].shape), tuple(imgs[s].shape)))
            lv[s] = SmoothLevel(flows[s].data_ptr(), imgs[s].data_ptr(), None, H, W)
        _alert_not_deterministic('uof_smooth_loss_fwd')
        sums, loss = torch.empty(S * B * 2 + _lib.SUMS_EXTRA, device=dev, dtype=torch.float32), torch.empty(B, device=dev, dtype=torch.float32)
        with torch.cuda.device_of(flows[0]):
            _lib.call('uof_smooth_loss_fwd', lv, S, B, Bimg, _p(sums), _p(loss), _stream(flows[0]))
        ctx.save_for_backward(*flows, *imgs)
        ctx.S = S
        return loss

    @staticmethod
    def backward(ctx, g):
        S = ctx.S
        flows, imgs = ctx.saved_tensors[:S], ctx.saved_tensors[S:]
        B, Bimg = flows[0].shape[0], imgs[0].shape[0]
        lv = _levels(SmoothLevel, S)
        gf = [torch.empty_like(t) for t in flows]
        for s in range(S):
            _, _, H, W = flows[s].shape
            lv[s] = SmoothLevel(flows[s].data_ptr(), imgs[s].data_ptr(), gf[s].data_ptr(), H, W)
        g = g.contiguous()
        with torch.cuda.device_of(g):
            _lib.call('uof_smooth_loss_bwd', lv, S, B, Bimg, _p(g), _stream(g))
        return (None, *gf, *([None] * S))


def flow_smooth_loss(optical_flows, img_pyramid, num_scales=3):
    """Drop-in for Model_flow.compute_loss_flow_smooth (model_flow_paper.py:168-177).  The flow batch
    may be a multiple of the image batch (sample b reads image b % B_img): both directions in one launch."""
    S = num_scales
    _require_cuda(*optical_flows[:S], *img_pyramid[:S])
    return _SmoothLoss.apply(S, *optical_flows[:S], *img_pyramid[:S])


# ------------------------------------------------------------------------------------------ a8
class _ConsisLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, *tensors):
        ff = [t.contiguous() for t in tensors[:S]]
        fb = [t.contiguous() for t in tensors[S:2 * S]]
        wf = [t.contiguous() for t in tensors[2 * S:3 * S]]
        B = ff[0].shape[0]
        dev = ff[0].device
        lv = _levels(ConsisLevel, S)
        for s in range(S):
            _, _, H, W = ff[s].shape
            # the kernel indexes flow_bwd and weight_fwd with flow_fwd's sizes: a weight map of another level would be an
            # out-of-bounds device read, so refuse what the reference would have broadcast or rejected
            if tuple(ff[s].shape) != (B, 2, H, W) or fb[s].shape != ff[s].shape or tuple(wf[s].shape) != (B, 1, H, W):
                raise ValueError('flow_consis_loss level %d: fwd flow %r, bwd flow %r must both be (B,2,H,W) and the weight %r (B,1,H,W)'
                                 % (s, tuple(ff[s].shape), tuple(fb[s].shape), tuple(wf[s].shape)))
            lv[s] = ConsisLevel(ff[s].data_ptr(), fb[s].data_ptr(), wf[s].data_ptr(), None, H, W)
        _alert_not_deterministic('uof_consis_loss_fwd')
        sums, loss = torch.empty(S * B * 2 + _lib.SUMS_EXTRA, device=dev, dtype=torch.float32), torch.empty(B, device=dev, dtype=torch.float32)
        with torch.cuda.device_of(ff[0]):
            _lib.call('uof_consis_loss_fwd', lv, S, B, _p(sums), _p(loss), _stream(ff[0]))
        ctx.save_for_backward(sums, *ff, *fb, *wf)
        ctx.S = S
        return loss

    @staticmethod
    def backward(ctx, g):
        S = ctx.S
        sums, *rest = ctx.saved_tensors
        ff, fb, wf = rest[:S], rest[S:2 * S], rest[2 * S:3 * S]
        B = ff[0].shape[0]
        lv = _levels(ConsisLevel, S)
        gf = [torch.empty_like(t) for t in ff]
        for s in range(S):
            _, _, H, W = ff[s].shape
            lv[s] = ConsisLevel(ff[s].data_ptr(), fb[s].data_ptr(), wf[s].data_ptr(), gf[s].data_ptr(), H, W)
        g = g.contiguous()
        with torch.cuda.device_of(g):
            _lib.call('uof_consis_loss_bwd', lv, S, B, _p(sums), _p(g), _stream(g))
        return (None, *gf, *([None] * (2 * S)))


def flow_consis_loss(fwd_flows, bwd_flows, weights_fwd, num_scales=3):
    """Drop-in for Model_flow.compute_loss_flow_consis (model_flow_paper.py:180-195); gradient reaches
    the forward flows only (bwd flows and the weight are detached in the reference)."""
    S = num_scales
    _require_cuda(*fwd_flows[:S], *bwd_flows[:S], *weights_fwd[:S])
    return _ConsisLoss.apply(S, *fwd_flows[:S], *[t.detach() for t in bwd_flows[:S]],
                             *[t.detach() for t in weights_fwd[:S]])


# ------------------------------------------------------------------------------------ a11 glue
def _sample_dense(g, C, H, W):
    """(B,C,H,W) view that is dense inside a sample (any batch stride): what torch.cat's backward hands out."""
    return g.stride(3) == 1 and g.stride(2) == W and g.stride(1) == H * W and g.stride(0) >= C * H * W


class _BiasLeakyReLU(torch.autograd.Function):
    """y <- lrelu(y + bias[c]) in place on a fresh convolution output; backward fuses LeakyReLU-backward with the bias
    gradient reduction (glue around the kept-PyTorch convolutions, net_utils.py:7-11).

    `fork=True` returns the activation TWICE (two tensors on the same storage), one for each of its two consumers in the
    decoder's dense block (next convolution and torch.cat, pwc_tf.py:119-131).  Autograd then delivers the two
    gradients separately and the backward kernel sums them on the fly -- channel slices of cat gradients included, through
    their batch stride -- instead of autograd's strided add kernel plus a .contiguous() copy per activation."""

    @staticmethod
    def forward(ctx, y, bias, slope, fork):
        assert y.is_contiguous()
        B, C, H, W = y.shape
        with torch.cuda.device_of(y):
            _lib.call('uof_bias_lrelu_fwd', _p(y), _p(bias), B, C, H, W, float(slope), _stream(y))
        ctx.mark_dirty(y)
        ctx.save_for_backward(y)
        ctx.slope = float(slope)
        if fork:
            return y, y.detach()
        return y

    @staticmethod
    def backward(ctx, g1, g2=None):
        (y,) = ctx.saved_tensors
        B, C, H, W = y.shape
        if g1 is None:
            g1, g2 = g2, None
        if g1 is None:
            return None, None, None, None
        if not _sample_dense(g1, C, H, W):
            g1 = g1.contiguous()
        if g2 is not None and not _sample_dense(g2, C, H, W):
            g2 = g2.contiguous()
        gx = torch.empty_like(y)
        gbias = torch.empty(C, device=y.device, dtype=torch.float32)
        _alert_not_deterministic('uof_bias_lrelu_bwd (bias gradient)')
        with torch.cuda.device_of(y):
            _lib.call('uof_bias_lrelu_bwd2', _p(g1), g1.stride(0), _p(g2) if g2 is not None else None,
                      g2.stride(0) if g2 is not None else 0, _p(y), _p(gx), _p(gbias), B, C, H, W, ctx.slope, _stream(y))
        return gx, gbias, None, None


def bias_leaky_relu_(y: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.1, fork: bool = False):
    """In-place lrelu(y + bias[None,:,None,None]); `y` must be a freshly produced (non-leaf) contiguous NCHW tensor.
    `fork=True` returns two tensors on the same storage for an activation with two consumers (see _BiasLeakyReLU)."""
    _require_cuda(y, bias)
    if not y.is_contiguous():
        y = y.contiguous()
    return _BiasLeakyReLU.apply(y, bias.contiguous(), negative_slope, bool(fork))


class _BiasLeakyReLUTo(torch.autograd.Function):
    """lrelu(y + bias) of a fresh convolution output, written straight to where its consumers read it: `dsts` is a list of
    one or two `(buffer, channel_offset)` pairs (buffer = a pre-allocated (B,Ctot,H,W) concat buffer, or `None` for "in
    place on y").  Returns one tensor per destination: the channel slice of the buffer (or y).  With `_CatAlias` this makes
    the dense block's concatenations (pwc_tf.py:113-118) copy-free.  The buffers are not autograd inputs -- they travel in
    a Python list -- so the returned slices are ordinary outputs of this node."""

    @staticmethod
    def forward(ctx, y, bias, slope, dsts):
        assert y.is_contiguous() and 1 <= len(dsts) <= 2
        B, C, H, W = y.shape
        outs, ptrs = [], []
        for buf, off in dsts:
            if buf is None:
                o = y
            else:
                assert buf.is_contiguous() and buf.shape[0] == B and tuple(buf.shape[2:]) == (H, W) and off + C <= buf.shape[1]
                o = buf[:, off:off + C]
            outs.append(o)
            ptrs += [_p(o), o.stride(0)]
        if len(dsts) == 1:
            ptrs += [None, 0]
        with torch.cuda.device_of(y):
            _lib.call('uof_bias_lrelu_fwd2', _p(y), _p(bias), *ptrs, B, C, H, W, float(slope), _stream(y))
        if dsts[0][0] is None or (len(dsts) == 2 and dsts[1][0] is None):
            ctx.mark_dirty(y)
        ctx.save_for_backward(outs[0])
        ctx.slope = float(slope)
        return tuple(o if o is y else o.detach() for o in outs) if len(outs) > 1 else (outs[0] if outs[0] is y else outs[0].detach())

    @staticmethod
    def backward(ctx, g1, g2=None):
        (y,) = ctx.saved_tensors
        B, C, H, W = y.shape
        if g1 is None:
            g1, g2 = g2, None
        if g1 is None:
            return None, None, None, None
        if not _sample_dense(g1, C, H, W):
            g1 = g1.contiguous()
        if g2 is not None and not _sample_dense(g2, C, H, W):
            g2 = g2.contiguous()
        gx = torch.empty((B, C, H, W), device=y.device, dtype=torch.float32)
        gbias = torch.empty(C, device=y.device, dtype=torch.float32)
        _alert_not_deterministic('uof_bias_lrelu_bwd (bias gradient)')
        with torch.cuda.device_of(y):
            _lib.call('uof_bias_lrelu_bwd3', _p(g1), g1.stride(0), _p(g2) if g2 is not None else None,
                      g2.stride(0) if g2 is not None else 0, _p(y), y.stride(0), _p(gx), _p(gbias), B, C, H, W, ctx.slope,
                      _stream(y))
        return gx, gbias, None, None


def bias_leaky_relu_to(y: torch.Tensor, bias: torch.Tensor, dsts, negative_slope: float = 0.1):
    """lrelu(y + bias[None,:,None,None]) written to one or two destinations (see _BiasLeakyReLUTo); `y` must be a freshly
    produced contiguous NCHW convolution output.  Returns a tuple with one tensor per destination."""
    _require_cuda(y, bias)
    if not y.is_contiguous():
        y = y.contiguous()
    out = _BiasLeakyReLUTo.apply(y, bias.contiguous(), negative_slope, list(dsts))
    return out if isinstance(out, tuple) else (out,)


class _CatAlias(torch.autograd.Function):
    """torch.cat(parts, 1) when the parts ALREADY are the consecutive channel slices of `buf` (written there by
    bias_leaky_relu_to): returns `buf` itself, no copy; backward hands each part its channel slice of the gradient (views)."""

    @staticmethod
    def forward(ctx, buf_box, *parts):
        buf = buf_box[0]
        off = 0
        for t in parts:
            assert t.data_ptr() == buf.data_ptr() + off * buf.stride(1) * buf.element_size() and t.stride(0) == buf.stride(0), \
                'cat_alias: a part is not the expected slice of the buffer'
            off += t.shape[1]
        assert off == buf.shape[1]
        ctx.sizes = [t.shape[1] for t in parts]
        return buf.detach()

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for c in ctx.sizes:
            outs.append(g[:, off:off + c])
            off += c
        return (None, *outs)


def cat_alias(buf: torch.Tensor, parts):
    """== torch.cat(parts, 1), given that `parts` are the channel slices of `buf` in order (no copy)."""
    return _CatAlias.apply([buf], *parts)


class _WeightedMeanSum(torch.autograd.Function):
    """sum_k w_k * mean(t_k): the objective of train.py:147-150 in one launch each way."""

    @staticmethod
    def forward(ctx, weights, *terms):
        ts = [t.contiguous() for t in terms]
        K = len(ts)
        ptrs = (ctypes.c_void_p * K)(*[t.data_ptr() for t in ts])
        w = (ctypes.c_float * K)(*[float(v) for v in weights])
        n = (ctypes.c_int * K)(*[t.numel() for t in ts])
        out = torch.empty((), device=ts[0].device, dtype=torch.float32)
        with torch.cuda.device_of(ts[0]):
            _lib.call('uof_weighted_mean_sum_fwd', ptrs, w, n, K, _p(out), _stream(ts[0]))
        ctx.meta = (tuple(float(v) for v in weights), tuple(tuple(t.shape) for t in ts))
        return out

    @staticmethod
    def backward(ctx, g):
        weights, shapes = ctx.meta
        K = len(shapes)
        g = g.contiguous()
        outs = [torch.empty(sh, device=g.device, dtype=torch.float32) for sh in shapes]
        ptrs = (ctypes.c_void_p * K)(*[t.data_ptr() for t in outs])
        w = (ctypes.c_float * K)(*weights)
        n = (ctypes.c_int * K)(*[t.numel() for t in outs])
        with torch.cuda.device_of(g):
            _lib.call('uof_weighted_mean_sum_bwd', _p(g), w, n, K, ptrs, _stream(g))
        return (None, *outs)


def weighted_mean_sum(terms, weights) -> torch.Tensor:
    """sum_k weights[k] * terms[k].mean() as a 0-d tensor (train.py:147-150: the weighted batch means of the loss pack)."""
    terms = list(terms)
    _require_cuda(*terms)
    if len(terms) != len(weights) or not 1 <= len(terms) <= 8:
        raise ValueError('weighted_mean_sum: need 1..8 terms and as many weights')
    return _WeightedMeanSum.apply(tuple(weights), *terms)


class _SplitAt(torch.autograd.Function):
    """x -> (x[:n], x[n:]) as views; the backward pass is ONE concatenation of the two gradients.  (`torch.split` + `torch.cat`
    of two of three parts costs a copy forward; two plain slices cost two zero-filled full-size tensors and an add backward.)"""

    @staticmethod
    def forward(ctx, x, n):
        ctx.n, ctx.shape = n, tuple(x.shape)
        return x[:n], x[n:]

    @staticmethod
    def backward(ctx, ga, gb):
        n, shape = ctx.n, ctx.shape
        if ga is None and gb is None:
            return None, None
        ref = ga if ga is not None else gb
        if ga is None:
            ga = ref.new_zeros((n,) + shape[1:])
        if gb is None:
            gb = ref.new_zeros((shape[0] - n,) + shape[1:])
        return torch.cat((ga, gb), 0), None


def split_at(x: torch.Tensor, n: int):
    """(x[:n], x[n:]) along the batch, both views of `x` (do not modify them in place)."""
    return _SplitAt.apply(x, int(n))


class _UpsampleScaled(torch.autograd.Function):
    """scale * F.interpolate(x, size, mode='bilinear') (align_corners=False) in one kernel each way."""

    @staticmethod
    def forward(ctx, x, H, W, scale):
        xc = x.contiguous()
        B, C, h, w = xc.shape
        out = torch.empty((B, C, H, W), device=xc.device, dtype=torch.float32)
        with torch.cuda.device_of(xc):
            _lib.call('uof_upsample_bilinear_fwd', _p(xc), _p(out), B * C, h, w, H, W, float(scale), _stream(xc))
        ctx.dims = (B, C, h, w, H, W, float(scale))
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, h, w, H, W, scale = ctx.dims
        g = g.contiguous()
        gin = torch.empty((B, C, h, w), device=g.device, dtype=torch.float32)
        with torch.cuda.device_of(g):
            _lib.call('uof_upsample_bilinear_bwd', _p(g), _p(gin), B * C, h, w, H, W, scale, _stream(g))
        return gin, None, None, None


def upsample_bilinear_scaled(x: torch.Tensor, size, scale: float = 1.0) -> torch.Tensor:
    """== F.interpolate(x, size, mode='bilinear') * scale (pwc_tf.py:119 ... 177; `size` = (H, W) >= x's), fused."""
    _require_cuda(x)
    H, W = int(size[0]), int(size[1])
    return _UpsampleScaled.apply(x, H, W, float(scale))


# ------------------------------------------------------------------------------------------ a9
def _pyramid_launch(base, nimg, stride_img, B, C, H, W, sb, sc, sh, num_pyramid):
    lv = [torch.empty((nimg, B, C, int(H / 2 ** s), int(W / 2 ** s)), device=base.device, dtype=torch.float32)
          for s in range(1, num_pyramid)]
    ptrs = (ctypes.c_void_p * len(lv))(*[t.data_ptr() for t in lv])
    with torch.cuda.device_of(base):
        _lib.call('uof_img_pyramid', _p(base), stride_img, sb, sc, sh, ptrs, num_pyramid, nimg, B, C, H, W, _stream(base))
    return lv


def img_pyramid(img: torch.Tensor, num_pyramid: int):
    """Drop-in for Model_flow.generate_img_pyramid (model_flow_paper.py:54-60): no gradient.  `img` may be
    a strided view (e.g. one image of the stacked triplet); only the last dimension must be dense."""
    _require_cuda(img)
    img = img.detach()
    if img.stride(3) != 1:
        img = img.contiguous()
    B, C, H, W = img.shape
    outs = [img]
    if num_pyramid > 1:
        lv = _pyramid_launch(img, 1, 0, B, C, H, W, img.stride(0), img.stride(1), img.stride(2), num_pyramid)
        outs += [t[0] for t in lv]
    return outs


def img_pyramid_triplet(inputs: torch.Tensor, num_pyramid: int):
    """Pyramids of the three vertically stacked images of a training sample (B,3,3H,W) (model_flow_paper.py:206-209,
    229-231) in ONE launch.  Returns (pyr_l, pyr_c, pyr_r, stacked) where stacked[s] is the dense (3,B,3,h,w) tensor of
    level s >= 1 (None for level 0, which is a view of the input)."""
    _require_cuda(inputs)
    x = inputs.detach()
    if x.stride(3) != 1:
        x = x.contiguous()
    B, C, H3, W = x.shape
    H = H3 // 3
    views = [x[:, :, k * H:(k + 1) * H] for k in range(3)]
    pyr = [[v] for v in views]
    stacked = [None]
    if num_pyramid > 1:
        lv = _pyramid_launch(x, 3, H * x.stride(2), B, C, H, W, x.stride(0), x.stride(1), x.stride(2), num_pyramid)
        for t in lv:
            stacked.append(t)
            for k in range(3):
                pyr[k].append(t[k])
    return pyr[0], pyr[1], pyr[2], stacked


TRIPLET_SLOTS = (0, 2, 1)        # input image (left, centre, right) -> slot in the stacked [left; right; centre] order


def img_pyramid_triplet_stacked(inputs: torch.Tensor, num_pyramid: int):
    """Like img_pyramid_triplet, but every level -- level 0 included, as a dense copy -- comes back as ONE (3,B,3,h,w)
    tensor ordered [left; right; centre]: `stacked[0].view(3B,3,H,W)` is the encoder's batch and `stacked[s][:2]` viewed as
    (2B,3,h,w) the [left; right] source of the image warps, with no torch.cat.  Returns None when the fast path does not
    apply (H or W not a multiple of 4, more than 3 levels, unaligned input)."""
    _require_cuda(inputs)
    x = inputs.detach()
    B, C, H3, W = x.shape
    H = H3 // 3
    if (x.stride(3) != 1 or H3 % 3 or H % 4 or W % 4 or not 2 <= num_pyramid <= 3 or x.data_ptr() % 16
            or any(st % 4 for st in (H * x.stride(2), x.stride(0), x.stride(1), x.stride(2)))):
        return None
    lv0 = torch.empty((3, B, C, H, W), device=x.device, dtype=torch.float32)
    lv = [torch.empty((3, B, C, int(H / 2 ** s), int(W / 2 ** s)), device=x.device, dtype=torch.float32) for s in range(1, num_pyramid)]
    ptrs = (ctypes.c_void_p * len(lv))(*[t.data_ptr() for t in lv])
    slots = (ctypes.c_int * 3)(*TRIPLET_SLOTS)
    with torch.cuda.device_of(x):
        _lib.call('uof_img_pyramid_stacked', _p(x), H * x.stride(2), x.stride(0), x.stride(1), x.stride(2), _p(lv0), slots, ptrs,
                  num_pyramid, 3, B, C, H, W, _stream(x))
    return [lv0] + lv


# -------------------------------------------------------------------------------------- a12/a13
class _Splat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, flow):
        fc = flow.contiguous()
        B, H, W, _ = fc.shape
        uc = u.contiguous() if u is not None else None
        C = uc.shape[3] if uc is not None else 1
        out = torch.empty((B, H, W, C), device=fc.device, dtype=torch.float32)
        _alert_not_deterministic('uof_splat_fwd')
        with torch.cuda.device_of(fc):
            _lib.call('uof_splat_fwd', _p(uc), _p(fc), _p(out), B, H, W, C, _stream(fc))
        ctx.save_for_backward(fc, *(() if uc is None else (uc,)))
        ctx.C = C
        return out

    @staticmethod
    def backward(ctx, gout):
        fc = ctx.saved_tensors[0]
        uc = ctx.saved_tensors[1] if len(ctx.saved_tensors) > 1 else None
        B, H, W, _ = fc.shape
        gu = torch.empty_like(uc) if (uc is not None and ctx.needs_input_grad[0]) else None
        gf = torch.empty_like(fc) if ctx.needs_input_grad[1] else None
        if gu is None and gf is None:
            return None, None
        with torch.cuda.device_of(fc):
            _lib.call('uof_splat_bwd', _p(gout.contiguous()), _p(uc), _p(fc), _p(gu), _p(gf), B, H, W, ctx.C, _stream(fc))
        return gu, gf


def transformerFwd(U: torch.Tensor, flo: torch.Tensor, out_size=None) -> torch.Tensor:
    """Bilinear forward splat, NHWC: out[b,yc,xc,:] += U[b,y,x,:]*w (absent from the reference)."""
    B, H, W, _ = flo.shape
    if out_size is not None and tuple(out_size) != (H, W):
        raise ValueError('transformerFwd: out_size %r must equal the flow size %r' % (tuple(out_size), (H, W)))
    _require_cuda(U, flo)
    return _Splat.apply(U, flo)


def range_map(flow_nhwc: torch.Tensor) -> torch.Tensor:
    """transformerFwd(ones, flow): how many source pixels land on each target pixel."""
    _require_cuda(flow_nhwc)
    return _Splat.apply(None, flow_nhwc)


def occlusion_mask(flow_nhwc: torch.Tensor) -> torch.Tensor:
    """clamp(range_map, 0, 1), no gradient."""
    with torch.no_grad():
        r = _Splat.apply(None, flow_nhwc)
        with torch.cuda.device_of(r):
            _lib.call('uof_clamp01', _p(r), r.numel(), _stream(r))
    return r


def splat_targets(flow_nhwc: torch.Tensor) -> torch.Tensor:
    """(B,H,W,4) int64 flat target indices of the four splat corners, -1 when out of bounds."""
    _require_cuda(flow_nhwc)
    fc = flow_nhwc.contiguous()
    B, H, W, _ = fc.shape
    idx = torch.empty((B, H, W, 4), device=fc.device, dtype=torch.int64)
    with torch.cuda.device_of(fc):
        _lib.call('uof_splat_targets', _p(fc), _p(idx), B, H, W, _stream(fc))
    return idx


def fb_consistency_mask(flow_fwd, flow_rev, alpha=3.0, beta=0.05, align_corners=None):
    """|f_fwd + warp(f_rev, f_fwd)| < max(alpha, beta*|f_fwd|) -> (B,1,H,W) in {0,1}; no gradient."""
    _require_cuda(flow_fwd, flow_rev)
    assert flow_fwd.shape == flow_rev.shape and flow_fwd.shape[1] == 2
    ac = DEFAULT_ALIGN_CORNERS if align_corners is None else bool(align_corners)
    ff, fr = flow_fwd.detach().contiguous(), flow_rev.detach().contiguous()
    B, _, H, W = ff.shape
    mask = torch.empty((B, 1, H, W), device=ff.device, dtype=torch.float32)
    with torch.cuda.device_of(ff):
        _lib.call('uof_fb_consistency_mask', _p(ff), _p(fr), _p(mask), B, H, W, float(alpha), float(beta), _coord_flags(ac), _stream(ff))
    return mask
