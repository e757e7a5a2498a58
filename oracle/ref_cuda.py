"""ctypes bindings for oracle/_ref/libref_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

libref_oracle.so is the reference's own six .cu files compiled UNMODIFIED for sm_100a by
`make -C oracle ref` (sources stay under /root/reference; only the built .so travels).
This module drives the reference's `extern "C"` launchers with torch CUDA tensors exactly
the way the reference's C glue does (file:line cited per function), so GPU tests can compare
the product kernels with "the reference's own ops on identical inputs".
Needs a GPU; importing it on a CPU-only box is fine, calling it is not.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        for n in ("Correlation_forward_cuda_kernel", "Correlation_backward_cuda_kernel",
                  "PSROIPoolForwardLauncher", "PSROIPoolBackwardLauncher", "ROIAlignForwardLaucher",
                  "ROIAlignBackwardLaucher", "ROIPoolForwardLaucher", "ROIPoolBackwardLaucher",
                  "BilinearSamplerBHWD_updateOutput_cuda_kernel",
                  "BilinearSamplerBHWD_updateGradInput_cuda_kernel"):
            getattr(_lib, n).restype = C.c_int
        _lib.nms_cuda_compute.restype = None
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t):
    assert t.is_cuda and t.is_contiguous()
    return t


def psroi_forward(feat, rois, scale, PH, PW, G, D):
    """psroi_pooling/src/psroi_pooling_cuda.c:9-42 + functions/psroi_pool.py:18-33."""
    _chk(feat), _chk(rois)
    B, Cc, H, W = feat.shape
    R = rois.shape[0]
    top = torch.zeros(R, D, PH, PW, device=feat.device)
    mapping = torch.zeros(R, D, PH, PW, dtype=torch.int32, device=feat.device)
    lib().PSROIPoolForwardLauncher(_p(feat), C.c_float(scale), R, H, W, Cc, PH, PW, _p(rois), G, D,
                                   _p(top), _p(mapping), _stream())
    return top, mapping


def psroi_backward(top_diff, mapping, rois, feat_shape, scale, PH, PW, D):
    """psroi_pooling_cuda.c:45-82 (note pooled_width before pooled_height)."""
    B, Cc, H, W = feat_shape
    g = torch.zeros(feat_shape, device=top_diff.device)
    lib().PSROIPoolBackwardLauncher(_p(_chk(top_diff)), _p(mapping), B, rois.shape[0], C.c_float(scale), Cc,
                                    H, W, PW, PH, D, _p(g), _p(_chk(rois)), _stream())
    return g


def nms(dets, thresh):
    """nms/nms_gpu.py:6-11 + src/nms_cuda.c:8-19 (synchronises internally, default stream)."""
    _chk(dets)
    N = dets.shape[0]
    keep = torch.zeros(N, 1, dtype=torch.int32, device=dets.device)
    num = torch.zeros(1, dtype=torch.int32, device=dets.device)
    torch.cuda.synchronize()
    lib().nms_cuda_compute(_p(keep), _p(num), _p(dets), N, dets.shape[1], C.c_float(thresh))
    torch.cuda.synchronize()
    return keep[: int(num.item())]


def correlation_shape(H, W, pad, k, md, s1, s2):
    import math
    kr = (k - 1) // 2
    br = kr + md
    r = md // s2
    return ((2 * r + 1) ** 2, int(math.ceil((H + 2 * pad - 2 * br) / s1)), int(math.ceil((W + 2 * pad - 2 * br) / s1)))


def correlation_forward(in1, in2, pad, k, md, s1, s2):
    """correlation/src/correlation_cuda.c:11-93: zero-filled padded NHWC scratch + output."""
    _chk(in1), _chk(in2)
    B, Cc, H, W = in1.shape
    oc, oh, ow = correlation_shape(H, W, pad, k, md, s1, s2)
    r1 = torch.zeros(B, H + 2 * pad, W + 2 * pad, Cc, device=in1.device)
    r2 = torch.zeros_like(r1)
    out = torch.zeros(B, oc, oh, ow, device=in1.device)
    ok = lib().Correlation_forward_cuda_kernel(
        _p(out), B, oc, oh, ow, *out.stride(),
        _p(in1), Cc, H, W, *in1.stride(),
        _p(in2), Cc, *in2.stride(),
        _p(r1), _p(r2), pad, k, md, s1, s2, 1, _stream())
    assert ok == 1
    return out


def correlation_backward(in1, in2, gout, pad, k, md, s1, s2, slack=0):
    """correlation_cuda.c:95-180.  `slack` extra floats are allocated behind each gradient
    so the reference's out-of-range writes for stride1 > 1 land in memory we own."""
    _chk(in1), _chk(in2), _chk(gout)
    B, Cc, H, W = in1.shape
    r1 = torch.zeros(B, H + 2 * pad, W + 2 * pad, Cc, device=in1.device)
    r2 = torch.zeros_like(r1)
    n = in1.numel()
    buf1 = torch.zeros(n + slack, device=in1.device)
    buf2 = torch.zeros(n + slack, device=in1.device)
    g1 = buf1[:n].view_as(in1)
    g2 = buf2[:n].view_as(in1)
    ok = lib().Correlation_backward_cuda_kernel(
        _p(gout), *gout.shape, *gout.stride(),
        _p(in1), Cc, H, W, *in1.stride(),
        _p(in2), *in2.stride(),
        _p(g1), *g1.stride(),
        _p(g2), Cc, *g2.stride(),
        _p(r1), _p(r2), pad, k, md, s1, s2, 1, _stream())
    assert ok == 1
    return g1, g2


def roi_align_forward(feat, rois, scale, AH, AW):
    """roi_align/src/roi_align_cuda.c:7-40."""
    _chk(feat), _chk(rois)
    B, Cc, H, W = feat.shape
    top = torch.zeros(rois.shape[0], Cc, AH, AW, device=feat.device)
    lib().ROIAlignForwardLaucher(_p(feat), C.c_float(scale), rois.shape[0], H, W, Cc, AH, AW, _p(rois), _p(top), _stream())
    return top


def roi_align_backward(top_diff, rois, feat_shape, scale, AH, AW):
    """roi_align_cuda.c:42-76."""
    B, Cc, H, W = feat_shape
    g = torch.zeros(feat_shape, device=top_diff.device)
    lib().ROIAlignBackwardLaucher(_p(_chk(top_diff)), C.c_float(scale), B, rois.shape[0], H, W, Cc, AH, AW,
                                  _p(_chk(rois)), _p(g), _stream())
    return g


def roi_pool_forward(feat, rois, scale, PH, PW):
    """roi_pooling/src/roi_pooling_cuda.c:7-47."""
    _chk(feat), _chk(rois)
    B, Cc, H, W = feat.shape
    top = torch.zeros(rois.shape[0], Cc, PH, PW, device=feat.device)
    arg = torch.zeros(rois.shape[0], Cc, PH, PW, dtype=torch.int32, device=feat.device)
    lib().ROIPoolForwardLaucher(_p(feat), C.c_float(scale), rois.shape[0], H, W, Cc, PH, PW, _p(rois), _p(top), _p(arg), _stream())
    return top, arg


def roi_pool_backward(top_diff, argmax, rois, feat_shape, scale, PH, PW):
    """roi_pooling_cuda.c:49-88."""
    B, Cc, H, W = feat_shape
    g = torch.zeros(feat_shape, device=top_diff.device)
    lib().ROIPoolBackwardLaucher(_p(_chk(top_diff)), C.c_float(scale), B, rois.shape[0], H, W, Cc, PH, PW,
                                 _p(_chk(rois)), _p(g), _p(argmax), _stream())
    return g


def roi_crop_forward(img, grid):
    """roi_crop/src/roi_crop_cuda.c:15-52 (argument order as the C glue passes it)."""
    _chk(img), _chk(grid)
    B, Cc, H, W = img.shape
    R, gh, gw, _ = grid.shape
    out = torch.zeros(R, Cc, gh, gw, device=img.device)
    ok = lib().BilinearSamplerBHWD_updateOutput_cuda_kernel(
        out.shape[1], out.shape[3], out.shape[2], out.shape[0], Cc, H, W, B,
        _p(img), *img.stride(),
        _p(grid), grid.stride(0), grid.stride(3), grid.stride(1), grid.stride(2),
        _p(out), *out.stride(), _stream())
    assert ok == 1
    return out


def roi_crop_backward(img, grid, gout):
    """roi_crop_cuda.c:54-105.  Returns (grad_image, grad_grid); grad_grid stays zero."""
    _chk(img), _chk(grid), _chk(gout)
    B, Cc, H, W = img.shape
    gi = torch.zeros_like(img)
    gg = torch.zeros_like(grid)
    ok = lib().BilinearSamplerBHWD_updateGradInput_cuda_kernel(
        gout.shape[1], gout.shape[3], gout.shape[2], gout.shape[0], Cc, H, W, B,
        _p(img), *img.stride(),
        _p(grid), grid.stride(0), grid.stride(3), grid.stride(1), grid.stride(2),
        _p(gi), *gi.stride(),
        _p(gg), gg.stride(0), gg.stride(3), gg.stride(1), gg.stride(2),
        _p(gout), *gout.stride(), _stream())
    assert ok == 1
    return gi, gg
