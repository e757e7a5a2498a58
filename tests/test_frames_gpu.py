"""Frame preparation on the device (csrc/frames.cu through the C-ABI) against the numpy oracle and the outputs of the
reference's own blob.py (tests/golden/frames_reference.npz).  Bar: bit-identical to the oracle (same roundings, no fused
multiply-adds); against OpenCV's own code 1e-6 of max |x| (north-star fp tolerance: 1e-4)."""
import numpy as np
import pytest
import torch

import common
from oracle import frames as oracle

pytestmark = pytest.mark.gpu
SMALL = ["down_cap", "down_nocap", "up", "flip", "half", "tall", "same"]


@pytest.fixture(scope="module")
def golden():
    return np.load(common.GOLDEN + "/frames_reference.npz")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", SMALL)
def test_small_frames_vs_reference_and_oracle(golden, name):
    from d2t_b200 import ops
    target, max_size, cap, flipped, scale = golden[name + "_args"]
    im = golden[name + "_im"]
    h, w = im.shape[:2]
    dh, dw, s = ops.frames_resized_shape(h, w, int(target), int(max_size), bool(cap))
    assert s == scale
    want = golden[name + "_out"]
    mine, _ = oracle.prep_im_for_blob(im[:, ::-1] if flipped else im, oracle.PIXEL_MEANS, int(target), int(max_size), bool(cap))
    nhwc = ops.frames_prep(cu(im[None]), s, bool(flipped), nhwc=True).cpu().numpy()[0]
    nchw = ops.frames_prep(cu(im[None]), s, bool(flipped)).cpu().numpy()[0]
    assert nhwc.shape == want.shape == (dh, dw, 3)
    np.testing.assert_array_equal(nhwc, mine)
    np.testing.assert_array_equal(nchw, mine.transpose(2, 0, 1))
    assert np.abs(nhwc - want).max() <= 1e-6 * np.abs(want).max()


def test_blob_padding_and_batch(golden):
    from d2t_b200 import ops
    ims = np.stack([common.make_frame(36, 64, s) for s in (21, 22, 23)])
    s = oracle.im_scale_for(36, 64, 30, 50, True)
    want, info = oracle.frames_to_blob(ims, 30, 50, cap=True)
    for bh, bw in [(28, 50), (32, 52), (29, 51), (40, 77)]:                      # float4 and scalar store paths
        out = torch.full((3, 3, bh, bw), 7.0, device="cuda")                     # stale contents must be overwritten
        ops.frames_prep(cu(ims), s, blob_hw=(bh, bw), out=out)
        out = out.cpu().numpy()
        np.testing.assert_array_equal(out[:, :, :28, :50], want)
        assert (out[:, :, 28:] == 0).all() and (out[:, :, :, 50:] == 0).all()
        hwc = ops.frames_prep(cu(ims), s, blob_hw=(bh, bw), nhwc=True).cpu().numpy()
        np.testing.assert_array_equal(hwc[:, :28, :50], want.transpose(0, 2, 3, 1))
        assert (hwc[:, 28:] == 0).all() and (hwc[:, :, 50:] == 0).all()
    data, im_info = ops.frames_to_blob(cu(ims), 30, 50, cap=True)
    np.testing.assert_array_equal(data.cpu().numpy(), want)
    np.testing.assert_array_equal(im_info.numpy(), info)


@pytest.mark.parametrize("name", ["vid_cap", "vid_nocap"])
def test_full_size_frame_vs_reference(golden, name):
    """720x1280 VID frame -> 562x1000 (eval loops) / 600x1067 (minibatch.py): sampled outputs of blob.py + cv2."""
    from d2t_b200 import ops
    h, w, target, max_size, cap, seed, scale, dh, dw = golden[name + "_args"]
    im = common.make_frame(int(h), int(w), int(seed))
    data, info = ops.frames_to_blob(cu(im[None]), int(target), int(max_size), cap=bool(cap))
    assert tuple(data.shape) == (1, 3, int(dh), int(dw)) and info[0].tolist() == [dh, dw, np.float32(scale)]
    ys, xs = golden[name + "_yx"]
    got = data[0].permute(1, 2, 0).cpu().numpy()[ys, xs]
    np.testing.assert_array_equal(got, golden[name + "_vals"])
    ipp = golden[name + "_vals_ipp"]
    assert np.abs(got - ipp).max() <= 2e-4 * np.abs(ipp).max()


def test_properties_at_full_size():
    """Size-independent checks on a batch of four 720x1280 frames (the bench's two frame pairs)."""
    from d2t_b200 import ops
    frames = np.stack([common.make_frame(720, 1280, 30 + i) for i in range(4)])
    d = cu(frames)
    # scale 1: every output is float32(double(u8) - mean), exactly
    one = ops.frames_prep(d, 1.0).cpu().numpy()
    want = (frames.astype(np.float64) - oracle.PIXEL_MEANS).astype(np.float32).transpose(0, 3, 1, 2)
    np.testing.assert_array_equal(one, want)
    # flipped=True equals preparing the mirrored frame (minibatch.py:77-78)
    s = oracle.im_scale_for(720, 1280, 600, 1000, True)
    a = ops.frames_prep(d, s, flipped=True)
    b = ops.frames_prep(cu(frames[:, :, ::-1]), s)
    assert torch.equal(a, b)
    # a constant frame stays constant (the two interpolation weights sum to one up to an ulp), frames are independent
    const = np.full((1, 720, 1280, 3), 200, np.uint8)
    c = ops.frames_prep(cu(const), s).cpu().numpy()
    for ch in range(3):
        v = np.float32(200.0 - oracle.PIXEL_MEANS[0, 0, ch])
        assert np.abs(c[0, ch] - v).max() <= 2e-5 * abs(v) + 1e-5
    single = ops.frames_prep(cu(frames[2:3]), s)
    assert torch.equal(ops.frames_prep(d, s)[2:3], single)
    # whole frame against the oracle, bit for bit
    want, _ = oracle.frames_to_blob(frames[2:3], 600, 1000, cap=True)
    np.testing.assert_array_equal(single.cpu().numpy(), want)


def test_blob_py_mirror(golden):
    """model/utils/blob.py: the reference's two function names and argument order."""
    from model.utils.blob import im_list_to_blob, prep_im_for_blob
    a, sa = prep_im_for_blob(golden["down_nocap_im"], oracle.PIXEL_MEANS, 30, 50)
    b, sb = prep_im_for_blob(cu(golden["tall_im"]), oracle.PIXEL_MEANS, 33, 60)
    wa, wsa = oracle.prep_im_for_blob(golden["down_nocap_im"], oracle.PIXEL_MEANS, 30, 50)
    wb, wsb = oracle.prep_im_for_blob(golden["tall_im"], oracle.PIXEL_MEANS, 33, 60)
    assert (sa, sb) == (wsa, wsb)
    np.testing.assert_array_equal(a.cpu().numpy(), wa)
    np.testing.assert_array_equal(b.cpu().numpy(), wb)
    np.testing.assert_array_equal(im_list_to_blob([a, b]).cpu().numpy(), oracle.im_list_to_blob([wa, wb]))
    assert np.abs(a.cpu().numpy() - golden["down_nocap_out"]).max() <= 1e-6 * np.abs(golden["down_nocap_out"]).max()
    with pytest.raises(ValueError):
        prep_im_for_blob(golden["tall_im"].astype(np.float32), oracle.PIXEL_MEANS, 33, 60)


def test_argument_errors():
    from d2t_b200 import ops
    from d2t_b200._lib import D2TError
    d = cu(common.make_frame(20, 30, 1)[None])
    with pytest.raises(ValueError):
        ops.frames_prep(d.cpu(), 1.0)
    with pytest.raises(ValueError):
        ops.frames_prep(d.float(), 1.0)
    with pytest.raises(D2TError):
        ops.frames_prep(d, 2.0, blob_hw=(10, 10))                                # blob smaller than the resized frame


def _split_formula_exact(m):
    """csrc/frames.cu means_split: is (u8 - hi) - lo == float32(double(u8) - mean) for every u8?"""
    for c in range(3):
        hi = np.rint(m[c] * 256) / 256
        lo = np.float32(m[c] - hi)
        i = np.arange(256)
        got = ((i.astype(np.float32) - np.float32(hi)).astype(np.float32) - lo).astype(np.float32)
        if not np.array_equal(got, (i.astype(np.float64) - m[c]).astype(np.float32)):
            return False
    return True


def test_other_pixel_means_table_and_table_free_paths():
    """the mean subtraction is float32(double(u8) - mean): table-free when the kernel's hi/lo split reproduces it for all
    256 values (the reference's PIXEL_MEANS do), a 768-entry table otherwise -- both bit-identical to the oracle"""
    from d2t_b200 import ops
    assert _split_formula_exact(oracle.PIXEL_MEANS.reshape(3))
    rng = np.random.RandomState(0)
    exact, table = None, None
    while exact is None or table is None:
        m = rng.uniform(0, 255, 3)
        if _split_formula_exact(m):
            exact = m if exact is None else exact
        else:
            table = m if table is None else table
    ims = np.stack([common.make_frame(40, 56, s) for s in (41, 42)])
    ims[0, :, :, 0] = np.arange(56, dtype=np.uint8)[None, :] * 4 + 3             # every byte value somewhere
    ims[1, :, :, 1] = 255 - np.arange(40, dtype=np.uint8)[:, None] * 6
    for means in (exact, table):
        for s in (1.0, oracle.im_scale_for(40, 56, 33, 1000, False)):
            want = np.stack([oracle.prep_im_for_blob(im, means.reshape(1, 1, 3), 33 if s != 1.0 else 40, 1000)[0] for im in ims])
            got = ops.frames_prep(cu(ims), s, pixel_means=means, nhwc=True).cpu().numpy()
            np.testing.assert_array_equal(got, want)


def test_video_pair_blobs():
    """online_tubes.py:582-606: consecutive frame pairs of one video, each frame prepared once."""
    from d2t_b200 import ops
    frames = np.stack([common.make_frame(45, 80, 70 + i) for i in range(5)])
    blobs = ops.VideoPairBlobs(cu(frames), 30, 50)
    assert len(blobs) == 4
    want, info = oracle.frames_to_blob(frames, 30, 50, cap=True)                # = _get_image_blob per frame (cap applies)
    for i in range(4):
        s = blobs[i]
        np.testing.assert_array_equal(s['data'].cpu().numpy(), want[i:i + 2])   # torch.cat([t0, t1]) of the reference
        np.testing.assert_array_equal(s['im_info'].cpu().numpy(), info[i:i + 2])
        assert s['frame_number'].view(-1).tolist() == [i, i + 1]
    im, im_info = blobs.batch(1, 3)
    assert im.is_contiguous() and tuple(im.shape) == (3, 2, 3) + want.shape[2:]
    for p in range(3):
        np.testing.assert_array_equal(im[p].cpu().numpy(), want[1 + p:3 + p])
        np.testing.assert_array_equal(im_info[p].cpu().numpy(), info[1 + p:3 + p])
    with pytest.raises(IndexError):
        blobs.batch(2, 3)
    with pytest.raises(IndexError):
        blobs[4]
