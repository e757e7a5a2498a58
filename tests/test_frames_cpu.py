"""Frame preparation, CPU side: the numpy restatement of blob.py + cv2.resize (oracle/frames.py)
against outputs of the reference's own blob.py (tests/golden/frames_reference.npz, made by make_golden_frames.py), and
the library's host-side shape arithmetic."""
import numpy as np
import pytest

import common
from oracle import frames as oracle

SMALL = ["down_cap", "down_nocap", "up", "flip", "half", "tall", "same"]


@pytest.fixture(scope="module")
def golden():
    return np.load(common.GOLDEN + "/frames_reference.npz")


def test_oracle_matches_reference_blob_py(golden):
    for name in SMALL:
        target, max_size, cap, flipped, scale = golden[name + "_args"]
        im = golden[name + "_im"]
        out, s = oracle.prep_im_for_blob(im[:, ::-1] if flipped else im, oracle.PIXEL_MEANS, int(target), int(max_size), bool(cap))
        want = golden[name + "_out"]
        assert s == scale and out.shape == want.shape, name
        # OpenCV's own code: identical up to the last bit of the two fp32 passes (its SIMD rows may fuse a multiply-add)
        assert np.abs(out - want).max() <= 1e-6 * np.abs(want).max(), name
        if name in ("down_cap", "down_nocap", "flip", "same"):
            np.testing.assert_array_equal(out, want)


def test_oracle_blob_padding(golden):
    a, _ = oracle.prep_im_for_blob(golden["down_cap_im"], oracle.PIXEL_MEANS, 30, 50, True)
    b, _ = oracle.prep_im_for_blob(golden["tall_im"], oracle.PIXEL_MEANS, 33, 60, True)
    blob = oracle.im_list_to_blob([a, b])
    want = golden["blob_pair"]
    assert blob.shape == want.shape
    assert np.abs(blob - want).max() <= 1e-6 * np.abs(want).max()
    assert (blob[0, a.shape[0]:] == 0).all() and (blob[1, :, b.shape[1]:] == 0).all()


@pytest.mark.parametrize("name", ["vid_cap", "vid_nocap"])
def test_oracle_full_size_frame(golden, name):
    h, w, target, max_size, cap, seed, scale, dh, dw = golden[name + "_args"]
    im = common.make_frame(int(h), int(w), int(seed))
    out, s = oracle.prep_im_for_blob(im, oracle.PIXEL_MEANS, int(target), int(max_size), bool(cap))
    assert s == scale and out.shape == (int(dh), int(dw), 3)
    ys, xs = golden[name + "_yx"]
    np.testing.assert_array_equal(out[ys, xs], golden[name + "_vals"])          # reductions: bit for bit
    # the IPP-dispatched OpenCV build of the container rounds its coordinates differently: 1e-4 of max |x| at 1000 px
    ipp = golden[name + "_vals_ipp"]
    assert np.abs(out[ys, xs] - ipp).max() <= 2e-4 * np.abs(ipp).max()


def test_library_shape_arithmetic_matches_reference():
    from d2t_b200 import ops
    for h, w, target, max_size in [(720, 1280, 600, 1000), (1280, 720, 600, 1000), (480, 640, 600, 1000), (333, 500, 600, 1000),
                                   (36, 64, 30, 50), (50, 21, 33, 60), (375, 1242, 600, 1000), (5, 7, 600, 1000)]:
        for cap in (False, True):
            s = oracle.im_scale_for(h, w, target, max_size, cap)
            dh, dw = oracle.resized_shape(h, w, s)
            assert ops.frames_resized_shape(h, w, target, max_size, cap) == (dh, dw, s)
