"""Torch-tensor front end of libd2t_b200.so: raw-pointer calls + autograd glue.

Every function takes contiguous fp32 CUDA tensors, enqueues on torch's current stream and
allocates its outputs / workspaces from torch's caching allocator (stream-ordered, no
cudaMalloc on the hot path).  The reference-shaped classes under ``model/`` are thin shells
over the functions here.  Nothing in this file computes on the CPU.
"""
import ctypes as C
import os

import torch

from ._lib import check, lib


# False routes every correlation through the fp32 SIMT kernels of csrc/correlation.cu
TENSOR_CORE_CORRELATION = True
# operand split of the tensor-core correlation: 16 = 3xFP16 (default), 3 = 3xTF32 (both fp32-accurate)
CORRELATION_PASSES = 16

# kernels of libd2t_b200.so enqueued since import (bench.py reports the delta as gpu_launches)
LAUNCHES = 0


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (d2t_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------- PSRoI
def psroi_mapping_channel(num_rois, pooled_h, pooled_w, group, out_dim, device):
    """The reference's `mappingchannel` output (psroi_pooling_kernel.cu:64-66, 77) is a pure
    function of the output index, c = (ctop*G + ph)*G + pw; built on demand instead of written by
    the kernel on every call (it would double the op's HBM writes)."""
    ctop = torch.arange(out_dim, device=device, dtype=torch.int32).view(1, out_dim, 1, 1)
    ph = torch.arange(pooled_h, device=device, dtype=torch.int32).view(1, 1, pooled_h, 1)
    pw = torch.arange(pooled_w, device=device, dtype=torch.int32).view(1, 1, 1, pooled_w)
    return ((ctop * group + ph) * group + pw).expand(num_rois, out_dim, pooled_h, pooled_w).contiguous()


def psroi_forward(features, rois, pooled_h, pooled_w, scale, group, out_dim, want_mapping=False):
    """psroi_pooling/functions/psroi_pool.py:18-33 -> d2t_psroi_forward.  Returns (top, mapping);
    mapping is None unless want_mapping (then the kernel writes it, as the reference's does).
    The library picks the table kernel (csrc/psroi.cu): per-plane fixed-point int32 tables when there is more than one
    (image, class, bin-row) item per SM, always for planes up to 64 wide (d2t_psroi_set_mode(0, 0) forces the exactly-rounded fp64 tables)."""
    _req(features, "features"), _req(rois, "rois")
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be [R, 5]")   # the reference returns 0 silently (psroi_pooling_cuda.c:17-20)
    B, Cc, H, W = features.shape
    R = rois.size(0)
    with torch.cuda.device_of(features):
        top = torch.empty(R, out_dim, pooled_h, pooled_w, device=features.device)
        mapping = torch.empty(R, out_dim, pooled_h, pooled_w, dtype=torch.int32, device=features.device) if want_mapping else None
        nbytes = lib().d2t_psroi_workspace_bytes(R, B, pooled_h, pooled_w)
        ws = _ws(nbytes, features.device)
        check(lib().d2t_psroi_forward(features.data_ptr(), B, Cc, H, W, rois.data_ptr(), R, scale, pooled_h, pooled_w,
                                      group, out_dim, top.data_ptr(), mapping.data_ptr() if want_mapping else None,
                                      ws.data_ptr(), ws.numel(), _stream()), "d2t_psroi_forward")
        _count(2)
    return top, mapping


def psroi_vote(features, rois, pooled_h, pooled_w, scale, group, out_dim, softmax=False):
    """Fused PSRoI pooling + 7x7 vote (+ softmax over the classes): rfcn.py:133-140 / 194-196 in one pass -- what
    ``AvgPool2d(7)(psroi(features, rois)).view(R, D)`` (then ``F.softmax(.., 1)``) returns, without the [R, D, 7, 7] tensor.
    -> [R, out_dim].  Eval-path operator (no autograd)."""
    _req(features, "features"), _req(rois, "rois")
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be [R, 5]")
    B, Cc, H, W = features.shape
    R = rois.size(0)
    with torch.cuda.device_of(features):
        vote = torch.empty(R, out_dim, device=features.device)
        ws = _ws(lib().d2t_psroi_vote_workspace_bytes(R, B, pooled_h, pooled_w, out_dim), features.device)
        check(lib().d2t_psroi_vote_forward(features.data_ptr(), B, Cc, H, W, rois.data_ptr(), R, scale, pooled_h, pooled_w,
                                           group, out_dim, int(softmax), vote.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _stream()), "d2t_psroi_vote_forward")
        _count(3)
    return vote


def psroi_backward(top_diff, rois, feature_size, pooled_h, pooled_w, scale, group, out_dim):
    _req(top_diff, "grad_output"), _req(rois, "rois")
    B, Cc, H, W = feature_size
    R = rois.size(0)
    with torch.cuda.device_of(top_diff):
        grad = torch.empty(B, Cc, H, W, device=top_diff.device)
        ws = _ws(lib().d2t_psroi_workspace_bytes(R, B, pooled_h, pooled_w), top_diff.device)
        check(lib().d2t_psroi_backward(top_diff.data_ptr(), B, Cc, H, W, rois.data_ptr(), R, scale, pooled_h, pooled_w,
                                       group, out_dim, grad.data_ptr(), 0, ws.data_ptr(), ws.numel(), _stream()),
              "d2t_psroi_backward")
        _count(3)                                   # psroi_prep, psroi_bwd_amax, psroi_bwd_limb
    return grad


def psroi_bins(rois, pooled_h, pooled_w, scale, height, width):
    _req(rois, "rois")
    bins = torch.empty(rois.size(0), pooled_h, pooled_w, 4, dtype=torch.int32, device=rois.device)
    with torch.cuda.device_of(rois):
        check(lib().d2t_psroi_bins(rois.data_ptr(), rois.size(0), scale, pooled_h, pooled_w, height, width,
                                   bins.data_ptr(), _stream()), "d2t_psroi_bins")
        _count(1)
    return bins


class _PSRoI(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ph, pw, scale, group, out_dim, holder):
        top, _ = psroi_forward(features, rois, ph, pw, scale, group, out_dim)
        ctx.cfg = (ph, pw, scale, group, out_dim, tuple(features.shape))
        ctx.save_for_backward(rois)
        if holder is not None:   # the reference Function keeps these on itself (psroi_pool.py:28-31)
            holder.output, holder.rois, holder.feature_size = top, rois, features.size()
        return top

    @staticmethod
    def backward(ctx, grad_top):
        ph, pw, scale, group, out_dim, fsize = ctx.cfg
        (rois,) = ctx.saved_tensors
        grad = psroi_backward(grad_top.contiguous(), rois, fsize, ph, pw, scale, group, out_dim)
        return grad, None, None, None, None, None, None, None


def psroi_pool(features, rois, ph, pw, scale, group, out_dim, holder=None):
    return _PSRoI.apply(features, rois, ph, pw, scale, group, out_dim, holder)


# ------------------------------------------------------------------------------- correlation
def correlation_shape(H, W, pad, k, md, s1, s2):
    out = (C.c_int * 3)()
    check(lib().d2t_correlation_shape(H, W, pad, k, md, s1, s2, out), "d2t_correlation_shape")
    return out[0], out[1], out[2]


_CORR_CACHE = {}      # (shape, pad, md, stride, passes, device) -> (x1, x2, CorrLayer); not for concurrent use from several threads


def correlation_forward(in1, in2, pad, k, md, s1, s2):
    _req(in1, "input1"), _req(in2, "input2")
    if in1.shape != in2.shape:
        raise ValueError("correlation inputs must have the same shape")
    B, Cc, H, W = in1.shape
    oc, oh, ow = correlation_shape(H, W, pad, k, md, s1, s2)
    if k == 1 and s1 == s2 and 1 <= md // s2 <= 8 and Cc >= 32 and TENSOR_CORE_CORRELATION:
        # every D&T configuration (rfcn.py:58-60): Gram tiles on the tensor cores (csrc/conv.cu, CORR mode),
        # fp32-accurate 3xFP16; the NCHW operands are first re-laid out as NHWC (+ each tensor's max |x| for its scale)
        from . import conv as dc
        with torch.cuda.device_of(in1):
            # plan + NHWC staging buffers are cached per (shape, parameters, device): a call is two re-layout launches (each
            # folds its tensor's max |x| in), the correlation, and a copy out of the plan's fixed output buffer
            key = (tuple(in1.shape), pad, md, s1, CORRELATION_PASSES, in1.device)
            ent = _CORR_CACHE.get(key)
            if ent is None:
                x1 = dc.ActTensor(B, H, W, Cc, cstride=(Cc + 31) // 32 * 32, device=in1.device, zero=False)
                x2 = dc.ActTensor(B, H, W, Cc, cstride=(Cc + 31) // 32 * 32, device=in1.device, zero=False)
                layer = dc.CorrLayer(x1, x2, pad, md, s1, passes=CORRELATION_PASSES, want_nchw=True)
                if len(_CORR_CACHE) >= 16:
                    _CORR_CACHE.clear()
                ent = _CORR_CACHE[key] = (x1, x2, layer)
            x1, x2, layer = ent
            x1.load_nchw_amax(in1)
            x2.load_nchw_amax(in2)
            return layer.run().clone()
    with torch.cuda.device_of(in1):
        out = torch.empty(B, oc, oh, ow, device=in1.device)
        check(lib().d2t_correlation_forward(in1.data_ptr(), in2.data_ptr(), B, Cc, H, W, pad, k, md, s1, s2,
                                            out.data_ptr(), _stream()), "d2t_correlation_forward")
        _count(1)
    return out


def _correlation_backward_tc(in1, in2, grad_out, pad, md, stride, need1, need2):
    """Every D&T configuration (kernel_size 1, stride1 == stride2, pad == max_displacement): the gradients w.r.t. both
    inputs as banded GEMMs on the tensor cores (csrc/conv.cu, CORRB; 3xFP16, fp32-accurate) -- 0.3 ms where the SIMT gather
    kernels take 12 ms on conv4 (bench.py ops.corr_conv4).  The NCHW operands are re-laid out as NHWC first; for stride 2
    the gradient lives on the even positions of the input (the correlation lattice), the rest is zero."""
    from . import conv as dc
    B, Cc, H, W = in1.shape
    r = md // stride
    Hl, Wl = -(-H // stride), -(-W // stride)
    x1, x2 = dc.ActTensor.from_nchw(in1), dc.ActTensor.from_nchw(in2)
    g = dc.ActTensor.from_nchw(grad_out)
    scratch = dc.CorrBwdScratch(*dc.CorrBwdScratch.need(B, Cc, Hl, Wl, r), device=in1.device)
    outs = []
    for which, need, other in ((1, need1, x2), (2, need2, x1)):
        if not need:
            outs.append(None)
            continue
        lat = dc.ActTensor(B, Hl, Wl, Cc, cstride=(Cc + 3) // 4 * 4, device=in1.device)
        dc.CorrBwdLayer(g, 0, other, lat, md, stride, which, scratch).run()
        low = lat.to_nchw()
        if stride == 1:
            outs.append(low)
        else:
            full = torch.zeros(B, Cc, H, W, device=in1.device)
            full[:, :, ::stride, ::stride] = low
            _count(2)
            outs.append(full)
    return outs[0], outs[1]


def correlation_backward(in1, in2, grad_out, pad, k, md, s1, s2, need1=True, need2=True):
    _req(in1, "input1"), _req(in2, "input2"), _req(grad_out, "grad_output")
    B, Cc, H, W = in1.shape
    if (k == 1 and s1 == s2 and pad == md and 1 <= md // s2 <= 8 and md % s2 == 0 and Cc >= 32 and Cc % 4 == 0
            and TENSOR_CORE_CORRELATION):
        with torch.cuda.device_of(in1):
            return _correlation_backward_tc(in1, in2, grad_out, pad, md, s1, need1, need2)
    with torch.cuda.device_of(in1):
        g1 = torch.empty_like(in1) if need1 else None
        g2 = torch.empty_like(in2) if need2 else None
        check(lib().d2t_correlation_backward(in1.data_ptr(), in2.data_ptr(), grad_out.data_ptr(), B, Cc, H, W, pad, k,
                                             md, s1, s2, g1.data_ptr() if need1 else None,
                                             g2.data_ptr() if need2 else None, _stream()), "d2t_correlation_backward")
        _count(int(need1) + int(need2))
    return g1, g2


class _Correlation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in1, in2, pad, k, md, s1, s2):
        ctx.cfg = (pad, k, md, s1, s2)
        ctx.save_for_backward(in1, in2)
        return correlation_forward(in1, in2, pad, k, md, s1, s2)

    @staticmethod
    def backward(ctx, grad_out):
        in1, in2 = ctx.saved_tensors
        g1, g2 = correlation_backward(in1, in2, grad_out.contiguous(), *ctx.cfg,
                                      need1=ctx.needs_input_grad[0], need2=ctx.needs_input_grad[1])
        return g1, g2, None, None, None, None, None


def correlation(in1, in2, pad, k, md, s1, s2):
    return _Correlation.apply(in1, in2, pad, k, md, s1, s2)


# ------------------------------------------------------------------------------- NMS
def nms_batched(dets, thresh, max_keep=0, n_valid=None):
    """dets [B, N, >=4] fp32, each list sorted by the caller.  Returns (keep [B, stride] int32,
    num_keep [B] int32), both on the device; no host synchronisation."""
    _req(dets, "dets")
    B, N, dim = dets.shape
    stride = max_keep if max_keep > 0 else max(N, 1)
    with torch.cuda.device_of(dets):
        keep = torch.empty(B, stride, dtype=torch.int32, device=dets.device)
        num = torch.empty(B, dtype=torch.int32, device=dets.device)
        ws = _ws(lib().d2t_nms_workspace_bytes(B, N), dets.device)
        check(lib().d2t_nms_batched(dets.data_ptr(), n_valid.data_ptr() if n_valid is not None else None, B, N, dim,
                                    thresh, max_keep, keep.data_ptr(), stride, num.data_ptr(), ws.data_ptr(),
                                    ws.numel(), _stream()), "d2t_nms_batched")
        _count(lib().d2t_nms_launch_count(N, max_keep))
    return keep, num


def nms(dets, thresh):
    """nms/nms_gpu.py:6-11: int32 [K, 1] kept indices.  K is data dependent, so like the
    reference (``keep[:num_out[0]]``) this reads one int back from the device."""
    if dets.shape[0] == 0:
        return dets.new_zeros((0, 1), dtype=torch.int32)
    keep, num = nms_batched(dets.unsqueeze(0), float(thresh))
    return keep[0, : int(num.item())].unsqueeze(1)


# ------------------------------------------------------------------------------- RoIAlign / RoIPool / RoICrop
def roi_align_forward(features, rois, ah, aw, scale):
    _req(features, "features"), _req(rois, "rois")
    B, Cc, H, W = features.shape
    with torch.cuda.device_of(features):
        top = torch.empty(rois.size(0), Cc, ah, aw, device=features.device)
        check(lib().ROIAlignForwardLaucher(features.data_ptr(), scale, rois.size(0), H, W, Cc, ah, aw, rois.data_ptr(),
                                           top.data_ptr(), _stream()), "ROIAlignForwardLaucher")
        _count(1)
    return top


# RoIAlign / RoIPool / RoICrop backward: True = 64-bit fixed-point accumulation (bit-identical from run to run; d2t_roi_*_backward_det),
# False = the reference's float atomics behind the reference-named launchers.  D2T_ROI_DETERMINISTIC=1 or set at run time.
DETERMINISTIC_ROI_BACKWARD = os.environ.get("D2T_ROI_DETERMINISTIC", "0") == "1"


def _roi_scratch(elems, device):
    return _ws(lib().d2t_roi_backward_scratch_bytes(elems), device)


def roi_align_backward(grad_top, rois, feature_size, ah, aw, scale, deterministic=None):
    _req(grad_top, "grad_output"), _req(rois, "rois")
    B, Cc, H, W = feature_size
    det = DETERMINISTIC_ROI_BACKWARD if deterministic is None else deterministic
    with torch.cuda.device_of(grad_top):
        grad = torch.zeros(B, Cc, H, W, device=grad_top.device)
        if det:
            sc = _roi_scratch(grad.numel(), grad.device)
            check(lib().d2t_roi_align_backward_det(grad_top.data_ptr(), scale, B, rois.size(0), H, W, Cc, ah, aw, rois.data_ptr(),
                                                   grad.data_ptr(), sc.data_ptr(), sc.numel(), _stream()),
                  "d2t_roi_align_backward_det")
            _count(3)
            return grad
        check(lib().ROIAlignBackwardLaucher(grad_top.data_ptr(), scale, B, rois.size(0), H, W, Cc, ah, aw,
                                            rois.data_ptr(), grad.data_ptr(), _stream()), "ROIAlignBackwardLaucher")
        _count(1)
    return grad


class _RoIAlign(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ah, aw, scale):
        ctx.cfg = (ah, aw, scale, tuple(features.shape))
        ctx.save_for_backward(rois)
        return roi_align_forward(features, rois, ah, aw, scale)

    @staticmethod
    def backward(ctx, grad_top):
        ah, aw, scale, fsize = ctx.cfg
        (rois,) = ctx.saved_tensors
        return roi_align_backward(grad_top.contiguous(), rois, fsize, ah, aw, scale), None, None, None, None


def roi_align(features, rois, ah, aw, scale):
    return _RoIAlign.apply(features, rois, ah, aw, scale)


def roi_pool_forward(features, rois, ph, pw, scale):
    _req(features, "features"), _req(rois, "rois")
    B, Cc, H, W = features.shape
    with torch.cuda.device_of(features):
        top = torch.empty(rois.size(0), Cc, ph, pw, device=features.device)
        argmax = torch.empty(rois.size(0), Cc, ph, pw, dtype=torch.int32, device=features.device)
        check(lib().ROIPoolForwardLaucher(features.data_ptr(), scale, rois.size(0), H, W, Cc, ph, pw, rois.data_ptr(),
                                          top.data_ptr(), argmax.data_ptr(), _stream()), "ROIPoolForwardLaucher")
        _count(1)
    return top, argmax


def roi_pool_backward(grad_top, argmax, rois, feature_size, ph, pw, scale, deterministic=None):
    _req(grad_top, "grad_output")
    B, Cc, H, W = feature_size
    det = DETERMINISTIC_ROI_BACKWARD if deterministic is None else deterministic
    with torch.cuda.device_of(grad_top):
        grad = torch.empty(B, Cc, H, W, device=grad_top.device)
        if det:
            sc = _roi_scratch(grad.numel(), grad.device)
            check(lib().d2t_roi_pool_backward_det(grad_top.data_ptr(), B, rois.size(0), H, W, Cc, ph, pw, grad.data_ptr(),
                                                  argmax.data_ptr(), sc.data_ptr(), sc.numel(), _stream()),
                  "d2t_roi_pool_backward_det")
            _count(3)
            return grad
        check(lib().ROIPoolBackwardLaucher(grad_top.data_ptr(), scale, B, rois.size(0), H, W, Cc, ph, pw,
                                           rois.data_ptr(), grad.data_ptr(), argmax.data_ptr(), _stream()),
              "ROIPoolBackwardLaucher")
        _count(1)
    return grad


class _RoIPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ph, pw, scale, holder):
        top, argmax = roi_pool_forward(features, rois, ph, pw, scale)
        ctx.cfg = (ph, pw, scale, tuple(features.shape))
        ctx.save_for_backward(rois, argmax)
        if holder is not None:
            holder.argmax, holder.rois, holder.feature_size = argmax, rois, features.size()
        return top

    @staticmethod
    def backward(ctx, grad_top):
        ph, pw, scale, fsize = ctx.cfg
        rois, argmax = ctx.saved_tensors
        return roi_pool_backward(grad_top.contiguous(), argmax, rois, fsize, ph, pw, scale), None, None, None, None, None


def roi_pool(features, rois, ph, pw, scale, holder=None):
    return _RoIPool.apply(features, rois, ph, pw, scale, holder)


def roi_crop_forward(images, grid):
    """images [B,C,H,W]; grid [R,gh,gw,2] in (y, x) order (roi_crop_cuda.c:15-52)."""
    _req(images, "input1"), _req(grid, "input2")
    B, Cc, H, W = images.shape
    R, gh, gw, two = grid.shape
    if two != 2 or R % B != 0:
        raise ValueError("grid must be [R, gh, gw, 2] with R a multiple of the image batch")
    with torch.cuda.device_of(images):
        out = torch.empty(R, Cc, gh, gw, device=images.device)
        check(lib().BilinearSamplerBHWD_updateOutput_cuda_kernel(
            Cc, gw, gh, R, Cc, H, W, B, images.data_ptr(), *images.stride(),
            grid.data_ptr(), grid.stride(0), grid.stride(3), grid.stride(1), grid.stride(2),
            out.data_ptr(), *out.stride(), _stream()), "BilinearSamplerBHWD_updateOutput_cuda_kernel")
        _count(1)
    return out


def roi_crop_backward(images, grid, grad_out, deterministic=None):
    _req(grid, "input2"), _req(grad_out, "grad_output")
    B, Cc, H, W = images.shape
    R, gh, gw, _ = grid.shape
    det = DETERMINISTIC_ROI_BACKWARD if deterministic is None else deterministic
    with torch.cuda.device_of(grad_out):
        gi = torch.zeros(B, Cc, H, W, device=grad_out.device)
        gg = torch.zeros_like(grid)   # never written by the reference either (roi_crop_cuda_kernel.cu:111-194)
        if det and grid.is_contiguous() and grad_out.is_contiguous():
            sc = _roi_scratch(gi.numel(), gi.device)
            check(lib().d2t_roi_crop_backward_det(grid.data_ptr(), grad_out.data_ptr(), gi.data_ptr(), B, Cc, H, W, R, gh, gw,
                                                  sc.data_ptr(), sc.numel(), _stream()), "d2t_roi_crop_backward_det")
            _count(3)
            return gi, gg
        check(lib().BilinearSamplerBHWD_updateGradInput_cuda_kernel(
            Cc, gw, gh, R, Cc, H, W, B, images.data_ptr(), *images.stride(),
            grid.data_ptr(), grid.stride(0), grid.stride(3), grid.stride(1), grid.stride(2),
            gi.data_ptr(), *gi.stride(),
            gg.data_ptr(), gg.stride(0), gg.stride(3), gg.stride(1), gg.stride(2),
            grad_out.data_ptr(), *grad_out.stride(), _stream()), "BilinearSamplerBHWD_updateGradInput_cuda_kernel")
        _count(1)
    return gi, gg


class _RoICrop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, grid):
        ctx.save_for_backward(images, grid)
        return roi_crop_forward(images, grid)

    @staticmethod
    def backward(ctx, grad_out):
        images, grid = ctx.saved_tensors
        return roi_crop_backward(images, grid, grad_out.contiguous())


def roi_crop(images, grid):
    return _RoICrop.apply(images, grid)


# ------------------------------------------------------------------------------- RPN proposal step
def proposal_decode(anchors, deltas, cls_prob, im_info, feat_stride):
    """-> boxes [B, H*W*A, 4], fg scores [B, H*W*A] in (y, x, a) order."""
    _req(anchors, "anchors"), _req(deltas, "bbox_deltas"), _req(cls_prob, "rpn_cls_prob"), _req(im_info, "im_info")
    B, A4, H, W = deltas.shape
    A = A4 // 4
    if cls_prob.shape != (B, 2 * A, H, W) or anchors.shape != (A, 4):
        raise ValueError("proposal_decode: inconsistent shapes")
    with torch.cuda.device_of(deltas):
        boxes = torch.empty(B, H * W * A, 4, device=deltas.device)
        scores = torch.empty(B, H * W * A, device=deltas.device)
        check(lib().d2t_proposal_decode(anchors.data_ptr(), A, deltas.data_ptr(), cls_prob.data_ptr(), im_info.data_ptr(),
                                        B, H, W, feat_stride, boxes.data_ptr(), scores.data_ptr(), _stream()),
              "d2t_proposal_decode")
        _count(1)
    return boxes, scores


def proposal_gather(boxes, scores, order, n_take):
    _req(boxes, "boxes"), _req(scores, "scores"), _req(order, "order", torch.int64)
    B, n_total, _ = boxes.shape
    with torch.cuda.device_of(boxes):
        dets = torch.empty(B, n_take, 5, device=boxes.device)
        check(lib().d2t_proposal_gather(boxes.data_ptr(), scores.data_ptr(), order.data_ptr(), B, n_total,
                                        order.size(1), n_take, dets.data_ptr(), _stream()), "d2t_proposal_gather")
        _count(1)
    return dets


# True: d2t_proposal_topk_gather (radix select + bitonic sort in ONE CTA per image, bit-identical order) instead of
# torch.sort (cub, device-wide) + d2t_proposal_gather.  Measured on B200 at 4 x 28728 scores -> top 6000: ~115 us against
# ~40 us for the cub kernels + gather -- four CTAs cannot match a device-wide radix sort -- so it is off by default.
HAND_WRITTEN_TOPK = os.environ.get("D2T_TOPK", "0") == "1"
# D2T_TOPK=2: the selection and sort in six short launches of our own (two 16-bit histogram levels + scans, compaction, a
# device-wide rank sort fused with the gather; d2t_proposal_topk_gather_split), bit-identical to the stable sort.  On B200 at
# 4 x 28728 -> 6000: 105 us against 118 us for torch.sort + gather alone (scripts/topk_bench.py), 1 % slower inside the step (5.49-5.51 against
# 5.44-5.47 ms: the n_take^2 rank sort holds the SMs the next conv kernel wants); at n_take = 12000 (training) the rank
# sort's quadratic cost loses outright (180 against 110 us).  Off by default.
SPLIT_TOPK = os.environ.get("D2T_TOPK", "0") == "2"


def proposal_topk_gather(boxes, scores, n_take, split=False):
    _req(boxes, "boxes"), _req(scores, "scores")
    B, n_total, _ = boxes.shape
    with torch.cuda.device_of(boxes):
        dets = torch.empty(B, n_take, 5, device=boxes.device)
        if split:      # two-level histogram select + compaction, then a device-wide rank sort fused with the gather
            sc = _ws(lib().d2t_proposal_topk_scratch_bytes(B, n_take), boxes.device)
            check(lib().d2t_proposal_topk_gather_split(boxes.data_ptr(), scores.data_ptr(), B, n_total, n_take, dets.data_ptr(),
                                                       sc.data_ptr(), sc.numel(), _stream()), "d2t_proposal_topk_gather_split")
            _count(6)                                   # hist, scan, hist, scan, compaction, rank sort + gather
            return dets
        check(lib().d2t_proposal_topk_gather(boxes.data_ptr(), scores.data_ptr(), B, n_total, n_take, dets.data_ptr(),
                                             _stream()), "d2t_proposal_topk_gather")
        _count(1)
    return dets


def proposal_write_rois(dets, keep, num_keep, post):
    B, n_take, _ = dets.shape
    with torch.cuda.device_of(dets):
        rois = torch.empty(B, post, 5, device=dets.device)
        check(lib().d2t_proposal_write_rois(dets.data_ptr(), keep.data_ptr(), keep.size(1), num_keep.data_ptr(), B,
                                            n_take, post, rois.data_ptr(), _stream()), "d2t_proposal_write_rois")
        _count(1)
    return rois


def proposals(anchors, deltas, cls_prob, im_info, feat_stride, pre_nms_topN, post_nms_topN, nms_thresh):
    """rpn/proposal_layer.py:67-159 for the whole batch without a python loop or a host sync:
    decode+clip kernel -> stable descending sort -> gather top pre_nms -> batched on-device NMS
    capped at post_nms -> padded rois [B, post, 5]."""
    boxes, scores = proposal_decode(anchors, deltas, cls_prob, im_info, feat_stride)
    n_total = scores.size(1)
    n_take = min(pre_nms_topN, n_total) if pre_nms_topN > 0 else n_total
    if SPLIT_TOPK and n_total <= 32768:
        dets = proposal_topk_gather(boxes, scores, n_take, split=True)
    elif HAND_WRITTEN_TOPK and lib().d2t_proposal_topk_supported(n_total, n_take):
        dets = proposal_topk_gather(boxes, scores, n_take)          # select + stable sort + gather, one launch
    else:
        order = torch.sort(scores, dim=1, descending=True, stable=True)[1]
        dets = proposal_gather(boxes, scores, order, n_take)
    keep, num = nms_batched(dets, float(nms_thresh), max_keep=post_nms_topN)
    return proposal_write_rois(dets, keep, num, post_nms_topN)


# ------------------------------------------------------------------------------- frame preparation
PIXEL_MEANS = (102.9801, 115.9465, 122.7717)        # config.py:257 (B, G, R)


def frames_resized_shape(src_h, src_w, target_size, max_size, cap):
    """(dst_h, dst_w, im_scale) of `prep_im_for_blob` (blob.py:43-50; cap=False) / the eval loops' `_get_image_blob`
    (demo.py:270-276; cap=True) -> d2t_frames_resized_shape (host arithmetic only)."""
    hw, s = (C.c_int * 2)(), C.c_double()
    check(lib().d2t_frames_resized_shape(int(src_h), int(src_w), int(target_size), int(max_size), int(bool(cap)), hw,
                                         C.byref(s)), "d2t_frames_resized_shape")
    return hw[0], hw[1], s.value


def frames_prep(frames, im_scale, flipped=False, pixel_means=PIXEL_MEANS, blob_hw=None, nhwc=False, out=None):
    """uint8 BGR frames [n, H, W, 3] on the device -> the float32 network input the reference builds on the host
    (blob.py:20-52, minibatch.py:77-78, roibatchLoader.py:183): [n, 3, blob_h, blob_w], or [n, blob_h, blob_w, 3] with
    nhwc=True; zero-padded to `blob_hw` (default: the resized size).  One launch of d2t_frames_prep."""
    _req(frames, "frames", torch.uint8)
    if frames.dim() != 4 or frames.size(3) != 3:
        raise ValueError("frames must be [n, H, W, 3] uint8 (BGR, as cv2.imread returns them)")
    n, H, W, _ = frames.shape
    dst_h, dst_w = int(round_half_even(H * im_scale)), int(round_half_even(W * im_scale))
    bh, bw = blob_hw if blob_hw is not None else (dst_h, dst_w)
    shape = (n, bh, bw, 3) if nhwc else (n, 3, bh, bw)
    with torch.cuda.device_of(frames):
        if out is None:
            out = torch.empty(shape, device=frames.device, dtype=torch.float32)
        elif tuple(out.shape) != shape:
            raise ValueError("out must be %s, got %s" % (shape, tuple(out.shape)))
        _req(out, "out")
        means = (C.c_double * 3)(*[float(m) for m in pixel_means])
        check(lib().d2t_frames_prep(frames.data_ptr(), n, H, W, means, float(im_scale), int(bool(flipped)), dst_h, dst_w,
                                    out.data_ptr(), bh, bw, int(bool(nhwc)), _stream()), "d2t_frames_prep")
    _count(1)
    return out


def round_half_even(v):
    """cvRound / np.round of a double (Python's round() is half-to-even as well)."""
    return round(float(v))


def frames_to_blob(frames, target_size=600, max_size=1000, cap=False, flipped=False, pixel_means=PIXEL_MEANS, out=None):
    """`_get_image_blob` (minibatch.py:58-88 with cap=False; demo.py:252-283 with cap=True) + the loader's NCHW permute
    for a batch of equally sized frames: (data [n, 3, h, w], im_info [n, 3] = (h, w, im_scale) on the host)."""
    n, H, W, _ = frames.shape
    dst_h, dst_w, s = frames_resized_shape(H, W, target_size, max_size, cap)
    data = frames_prep(frames, s, flipped, pixel_means, out=out)
    im_info = torch.tensor([[dst_h, dst_w, s]] * n, dtype=torch.float32)
    return data, im_info


class VideoPairBlobs(object):
    """`VideoDataset._create_video_blob` (online_tubes.py:582-606 / tracking_utils.py:420-447) on the device: the
    n - 1 consecutive frame pairs (t, t + 1) of one video.  The reference reads, prepares and uploads every inner frame
    twice (as t1 of pair i - 1 and as t0 of pair i); here all n frames are prepared by ONE d2t_frames_prep launch and
    sample i is the zero-copy window `data[i:i + 2]` of that blob.

    `blobs[i]` -> {'data': [2, 3, h, w], 'im_info': [2, 3], 'frame_number': [2, 1, 1]} as the reference's sample dict;
    `blobs.batch(first, pairs)` -> (im_data [pairs, 2, 3, h, w] contiguous, im_info [pairs, 2, 3]) for the engines."""

    def __init__(self, frames, target_size=600, max_size=1000, pixel_means=PIXEL_MEANS):
        self.data, info = frames_to_blob(frames, target_size, max_size, cap=True, pixel_means=pixel_means)
        self.im_info = info.to(self.data.device)
        n = self.data.size(0)
        self.frame_number = torch.arange(n, device=self.data.device).view(n, 1, 1)

    def __len__(self):
        return self.data.size(0) - 1

    def __getitem__(self, i):
        if not 0 <= i < len(self):
            raise IndexError(i)
        return {'data': self.data[i:i + 2], 'im_info': self.im_info[i:i + 2], 'frame_number': self.frame_number[i:i + 2]}

    def batch(self, first, pairs):
        if first < 0 or pairs < 1 or first + pairs > len(self):
            raise IndexError("pairs [%d, %d) of %d" % (first, first + pairs, len(self)))
        d = self.data
        s = d.stride()
        im = d.as_strided((pairs, 2) + tuple(d.shape[1:]), (s[0], s[0]) + tuple(s[1:]), d.storage_offset() + first * s[0])
        i = self.im_info
        info = i.as_strided((pairs, 2, 3), (i.stride(0), i.stride(0), 1), first * i.stride(0))
        return im.contiguous(), info.contiguous()
