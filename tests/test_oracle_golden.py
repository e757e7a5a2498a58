"""Pins the CPU oracle (oracle/oracle_cpu.c) against outputs of the reference itself:
  * tests/golden/ref_cuda.npz  -- the reference's own CUDA kernels run on a B200
    (tests/golden/make_golden_gpu.py through oracle/_ref/libref_oracle.so);
  * tests/golden/rpn_reference.npz -- the reference's own Python RPN code run on CPU
    (tests/golden/make_golden_rpn.py).
No GPU needed."""
import numpy as np

import cases
import common


def test_psroi_forward_backward(oracle, golden_cuda):
    for name, c in cases.psroi_cases().items():
        top, mapping = oracle.psroi_forward(c["feat"], c["rois"], c["scale"], c["P"], c["P"], c["G"], c["D"])
        # same bins + same row-major summation order + IEEE division => bit-exact
        np.testing.assert_array_equal(top, golden_cuda["psroi_%s_top" % name])
        np.testing.assert_array_equal(mapping, golden_cuda["psroi_%s_map" % name])
        g = oracle.psroi_backward(c["gtop"], c["rois"], c["feat"].shape, c["scale"], c["P"], c["P"], c["G"], c["D"])
        np.testing.assert_allclose(g, golden_cuda["psroi_%s_grad" % name], rtol=1e-5, atol=1e-6)  # atomics order


def test_nms_keep_sets_bit_exact(oracle, golden_cuda):
    for name, (dets, thresh) in cases.nms_cases().items():
        keep = oracle.nms(dets, thresh)
        np.testing.assert_array_equal(keep, golden_cuda["nms_%s" % name], err_msg=name)


def _rogue_mask(shape, s1):
    """Elements the reference's Correlation_backward_input2 hits with out-of-plane writes when
    stride1 > 1 (blockIdx*stride1 beyond the plane, correlation_cuda_kernel.cu:212-213,285):
    undefined there (racing stores), so excluded from the comparison."""
    B, C, H, W = shape
    mask = np.zeros(B * C * H * W + 4 * H * W, bool)
    if s1 > 1:
        by, bx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        off = (by * s1 * W + bx * s1)[(by * s1 >= H) | (bx * s1 >= W)]
        for nc in range(B * C):
            mask[nc * H * W + off] = True
    return mask[: B * C * H * W].reshape(shape)


def test_correlation_forward_backward(oracle, golden_cuda):
    for name, c in cases.corr_cases().items():
        p = c["params"]
        out = oracle.correlation_forward(c["in1"], c["in2"], *p)
        ref = golden_cuda["corr_%s_out" % name]
        assert out.shape == ref.shape
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6, err_msg=name)
        go = common.randn(out.shape, c["gseed"])
        g1, g2, oob = oracle.correlation_backward_ref(c["in1"], c["in2"], go, *p)
        np.testing.assert_allclose(g1, golden_cuda["corr_%s_g1" % name], rtol=1e-4, atol=1e-6, err_msg=name)
        ok = ~_rogue_mask(g2.shape, p[3])
        np.testing.assert_allclose(g2[ok], golden_cuda["corr_%s_g2" % name][ok], rtol=1e-4, atol=1e-6, err_msg=name)
        if p[3] == 1:
            assert oob == (0, 0)
        # where the reference is well defined it IS the adjoint of its forward
        t1, t2 = oracle.correlation_backward_true(c["in1"], c["in2"], go, *p)
        if p[0] == p[2] and p[1] == 1:   # pad == max_displacement, kernel_size 1 (every D&T configuration)
            np.testing.assert_allclose(g1, t1, rtol=1e-4, atol=1e-6, err_msg=name)
            np.testing.assert_allclose(g2[ok], t2[ok], rtol=1e-4, atol=1e-6, err_msg=name)


def test_roi_align_pool_crop(oracle, golden_cuda):
    c = cases.roi_cases()
    feat, rois, grid, scale = c["feat"], c["rois"], c["grid"], c["scale"]
    for ah in (7, 8):
        top = oracle.roi_align_forward(feat, rois, scale, ah, ah)
        np.testing.assert_allclose(top, golden_cuda["align%d_top" % ah], rtol=1e-6, atol=1e-7)
        g = oracle.roi_align_backward(common.randn(top.shape, 80 + ah), rois, feat.shape, scale, ah, ah)
        np.testing.assert_allclose(g, golden_cuda["align%d_grad" % ah], rtol=1e-5, atol=1e-6)
    top, arg = oracle.roi_pool_forward(feat, rois, scale, 7, 7)
    np.testing.assert_array_equal(top, golden_cuda["pool_top"])
    np.testing.assert_array_equal(arg, golden_cuda["pool_arg"])
    g = oracle.roi_pool_backward(common.randn(top.shape, 90), arg, feat.shape)
    np.testing.assert_allclose(g, golden_cuda["pool_grad"], rtol=1e-5, atol=1e-6)
    out = oracle.roi_crop_forward(feat, grid)
    np.testing.assert_allclose(out, golden_cuda["crop_out"], rtol=1e-5, atol=1e-6)
    gi = oracle.roi_crop_backward(common.randn(out.shape, 91), grid, feat.shape)
    np.testing.assert_allclose(gi, golden_cuda["crop_gimg"], rtol=1e-5, atol=1e-6)
    assert float(golden_cuda["crop_ggrid_absmax"]) == 0.0


def test_anchors_match_reference(oracle, golden_rpn):
    np.testing.assert_array_equal(oracle.generate_anchors(), golden_rpn["anchors_default"])
    np.testing.assert_array_equal(oracle.generate_anchors(scales=(4, 8, 16, 32)), golden_rpn["anchors_d2t"])
    # known-answer table of generate_anchors.py:19-37 is these values + 1 (MATLAB indexing)
    assert golden_rpn["anchors_default"][0].tolist() == [-84.0, -40.0, 99.0, 55.0]


def test_proposal_layer_matches_reference(oracle, golden_rpn):
    anchors = oracle.generate_anchors(scales=(4, 8, 16, 32)).astype(np.float32)
    cases_ = {"small": dict(B=2, H=10, W=14, seed=30), "config1": dict(B=1, H=19, W=32, seed=31, im_h=300, im_w=500),
              "full": dict(B=1, H=38, W=63, seed=32, im_h=600, im_w=1000)}
    for name, kw in cases_.items():
        prob, deltas, im_info = common.make_rpn_inputs(**kw)
        for key, pre, post in (("TEST", 6000, 300), ("TRAIN", 12000, 2000)):
            rois = oracle.proposal_layer(prob, deltas, im_info, anchors, pre, post, 0.7)
            ref = golden_rpn["rois_%s_%s" % (name, key)]
            # torch's exp/ops on CPU may differ from libm in the last ulp -> tolerance, same rows
            np.testing.assert_allclose(rois, ref, rtol=1e-5, atol=1e-3, err_msg="%s %s" % (name, key))


def test_decode_clip_matches_reference(oracle, golden_rpn):
    boxes, deltas = golden_rpn["bti_boxes"], golden_rpn["bti_deltas"]
    # express as a 1x1 feature map with 64 "anchors"
    A = boxes.shape[1]
    d = deltas[0].reshape(A * 4, 1, 1)[None]
    pred = oracle.proposal_decode(boxes[0], d, np.array([[600., 1000., 1.]], np.float32), stride=16)
    np.testing.assert_allclose(pred, golden_rpn["bti_clipped"], rtol=1e-6, atol=1e-4)
