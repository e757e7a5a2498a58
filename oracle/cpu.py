"""ctypes bindings for oracle/liboracle_cpu.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  All arrays are numpy float32 / int32, C-contiguous, NCHW.
The C side (oracle_cpu.c) cites the reference lines each function restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_cpu.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "cpu"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_iou.restype = C.c_float
        _lib.oracle_nms.restype = C.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


# ---------------------------------------------------------------- PSRoI
def psroi_forward(feat, rois, scale, PH, PW, G, D, contract=1, want_bins=False):
    feat, pf = _f(feat)
    rois, pr = _f(rois)
    B, Cc, H, W = feat.shape
    R = rois.shape[0]
    top = np.zeros((R, D, PH, PW), np.float32)
    mapping = np.zeros((R, D, PH, PW), np.int32)
    bins = np.zeros((R, PH, PW, 4), np.int32) if want_bins else None
    lib().oracle_psroi_forward(pf, B, Cc, H, W, pr, R, C.c_float(scale), PH, PW, G, D, contract,
                               top.ctypes.data_as(C.POINTER(C.c_float)),
                               _i(mapping), _i(bins) if want_bins else None)
    return (top, mapping, bins) if want_bins else (top, mapping)


def psroi_backward(top_diff, rois, feat_shape, scale, PH, PW, G, D, contract=1):
    top_diff, pt = _f(top_diff)
    rois, pr = _f(rois)
    B, Cc, H, W = feat_shape
    out = np.zeros(feat_shape, np.float32)
    lib().oracle_psroi_backward(pt, B, Cc, H, W, pr, rois.shape[0], C.c_float(scale), PH, PW, G, D,
                                contract, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


# ---------------------------------------------------------------- NMS
def iou(a, b, contract=1):
    a, pa = _f(a)
    b, pb = _f(b)
    return float(lib().oracle_iou(pa, pb, contract))


def nms(dets, thresh, contract=1, max_keep=0):
    """dets [N, >=4] sorted by the caller; returns int32 [K] kept indices (ascending)."""
    dets, pd = _f(dets)
    N = dets.shape[0]
    keep = np.zeros((max(N, 1),), np.int32)
    k = lib().oracle_nms(pd, N, dets.shape[1] if N else 5, C.c_float(thresh), contract, max_keep, _i(keep))
    return keep[:k].copy()


# ---------------------------------------------------------------- correlation
def correlation_shape(H, W, pad, k, md, s1, s2):
    o = (C.c_int * 3)()
    lib().oracle_correlation_shape(H, W, pad, k, md, s1, s2, o)
    return o[0], o[1], o[2]


def correlation_forward(in1, in2, pad, k, md, s1, s2):
    in1, p1 = _f(in1)
    in2, p2 = _f(in2)
    B, Cc, H, W = in1.shape
    oc, oh, ow = correlation_shape(H, W, pad, k, md, s1, s2)
    out = np.zeros((B, oc, oh, ow), np.float32)
    lib().oracle_correlation_forward(p1, p2, B, Cc, H, W, pad, k, md, s1, s2,
                                     out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def correlation_backward_ref(in1, in2, gout, pad, k, md, s1, s2):
    """Faithful restatement of the reference backward kernels.  Returns (g1, g2, oob) where
    oob = number of writes the reference would have made past the end of each tensor."""
    in1, p1 = _f(in1)
    in2, p2 = _f(in2)
    gout, pg = _f(gout)
    B, Cc, H, W = in1.shape
    g1 = np.zeros(in1.shape, np.float32)
    g2 = np.zeros(in1.shape, np.float32)
    oob = (C.c_int64 * 2)()
    lib().oracle_correlation_backward_ref(p1, p2, pg, B, Cc, H, W, pad, k, md, s1, s2,
                                          g1.ctypes.data_as(C.POINTER(C.c_float)),
                                          g2.ctypes.data_as(C.POINTER(C.c_float)),
                                          C.c_size_t(g1.size), oob)
    return g1, g2, (int(oob[0]), int(oob[1]))


def correlation_backward_true(in1, in2, gout, pad, k, md, s1, s2):
    in1, p1 = _f(in1)
    in2, p2 = _f(in2)
    gout, pg = _f(gout)
    B, Cc, H, W = in1.shape
    g1 = np.zeros(in1.shape, np.float32)
    g2 = np.zeros(in1.shape, np.float32)
    lib().oracle_correlation_backward_true(p1, p2, pg, B, Cc, H, W, pad, k, md, s1, s2,
                                           g1.ctypes.data_as(C.POINTER(C.c_float)),
                                           g2.ctypes.data_as(C.POINTER(C.c_float)))
    return g1, g2


# ---------------------------------------------------------------- RoIAlign / RoIPool / RoICrop
def roi_align_forward(feat, rois, scale, AH, AW):
    feat, pf = _f(feat)
    rois, pr = _f(rois)
    B, Cc, H, W = feat.shape
    top = np.zeros((rois.shape[0], Cc, AH, AW), np.float32)
    lib().oracle_roi_align_forward(pf, B, Cc, H, W, pr, rois.shape[0], C.c_float(scale), AH, AW,
                                   top.ctypes.data_as(C.POINTER(C.c_float)))
    return top


def roi_align_backward(top_diff, rois, feat_shape, scale, AH, AW):
    top_diff, pt = _f(top_diff)
    rois, pr = _f(rois)
    B, Cc, H, W = feat_shape
    out = np.zeros(feat_shape, np.float32)
    lib().oracle_roi_align_backward(pt, B, Cc, H, W, pr, rois.shape[0], C.c_float(scale), AH, AW,
                                    out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def roi_pool_forward(feat, rois, scale, PH, PW):
    feat, pf = _f(feat)
    rois, pr = _f(rois)
    B, Cc, H, W = feat.shape
    top = np.zeros((rois.shape[0], Cc, PH, PW), np.float32)
    arg = np.zeros((rois.shape[0], Cc, PH, PW), np.int32)
    lib().oracle_roi_pool_forward(pf, B, Cc, H, W, pr, rois.shape[0], C.c_float(scale), PH, PW,
                                  top.ctypes.data_as(C.POINTER(C.c_float)), _i(arg))
    return top, arg


def roi_pool_backward(top_diff, argmax, feat_shape):
    top_diff, pt = _f(top_diff)
    argmax = np.ascontiguousarray(argmax, np.int32)
    B, Cc, H, W = feat_shape
    R, _, PH, PW = top_diff.shape
    out = np.zeros(feat_shape, np.float32)
    lib().oracle_roi_pool_backward(pt, _i(argmax), B, Cc, H, W, R, PH, PW,
                                   out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def roi_crop_forward(img, grid):
    img, pi = _f(img)
    grid, pg = _f(grid)
    B, Cc, H, W = img.shape
    R, gh, gw, _ = grid.shape
    out = np.zeros((R, Cc, gh, gw), np.float32)
    lib().oracle_roi_crop_forward(pi, B, Cc, H, W, pg, R, gh, gw,
                                  out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def roi_crop_backward(gout, grid, img_shape):
    gout, po = _f(gout)
    grid, pg = _f(grid)
    B, Cc, H, W = img_shape
    R, gh, gw, _ = grid.shape
    out = np.zeros(img_shape, np.float32)
    lib().oracle_roi_crop_backward(po, B, Cc, H, W, pg, R, gh, gw,
                                   out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


# ---------------------------------------------------------------- RPN proposal step
def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    """rpn/generate_anchors.py:45-104 restated (float64 numpy, same op order)."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)

    def whctrs(a):
        w = a[2] - a[0] + 1
        h = a[3] - a[1] + 1
        return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)

    def mk(ws, hs, xc, yc):
        ws = ws[:, None]
        hs = hs[:, None]
        return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))

    base = np.array([1, 1, base_size, base_size], dtype=np.float64) - 1
    w, h, xc, yc = whctrs(base)
    size_ratios = (w * h) / ratios
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    ra = mk(ws, hs, xc, yc)
    out = []
    for i in range(ra.shape[0]):
        w, h, xc, yc = whctrs(ra[i])
        out.append(mk(w * scales, h * scales, xc, yc))
    return np.vstack(out)


def proposal_decode(anchors, deltas, im_info, stride=16):
    anchors, pa = _f(anchors)
    deltas, pd = _f(deltas)
    im_info, pi = _f(im_info)
    B, A4, H, W = deltas.shape
    A = A4 // 4
    boxes = np.zeros((B, H * W * A, 4), np.float32)
    lib().oracle_proposal_decode(pa, A, pd, pi, B, H, W, stride, boxes.ctypes.data_as(C.POINTER(C.c_float)))
    return boxes


def proposal_layer(cls_prob, deltas, im_info, anchors, pre_nms_topN, post_nms_topN, nms_thresh, stride=16):
    """rpn/proposal_layer.py:67-159 restated.  cls_prob [B, 2A, H, W] (fg = second half).
    Sort is a STABLE descending sort ('Sort tie-breaking')."""
    cls_prob = np.ascontiguousarray(cls_prob, np.float32)
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    scores = cls_prob[:, A:].transpose(0, 2, 3, 1).reshape(B, -1)
    boxes = proposal_decode(anchors, deltas, im_info, stride)
    out = np.zeros((B, post_nms_topN, 5), np.float32)
    for i in range(B):
        order = np.argsort(-scores[i], kind="stable")
        if 0 < pre_nms_topN < scores.size:
            order = order[:pre_nms_topN]
        p = boxes[i][order]
        s = scores[i][order][:, None]
        keep = nms(np.concatenate([p, s], 1), nms_thresh)
        if post_nms_topN > 0:
            keep = keep[:post_nms_topN]
        out[i, :, 0] = i
        out[i, : len(keep), 1:] = p[keep]
    return out
