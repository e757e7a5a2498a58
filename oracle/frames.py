"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's frame preparation --
`prep_im_for_blob` / `im_list_to_blob` (`lib/model/utils/blob.py:20-52`) as called by `_get_image_blob`
(`lib/roi_data_layer/minibatch.py:58-88`, training: no MAX_SIZE cap, optional horizontal flip) and by the eval loops
(`demo.py:252-283`, `lib/model/utils/online_tubes.py:640-670`: capped at TEST.MAX_SIZE), followed by the loader's
`permute(0, 3, 1, 2)` (`lib/roi_data_layer/roibatchLoader.py:183`).

The bilinear resize is OpenCV's `cv2.resize(..., fx, fy, INTER_LINEAR)` on float32 data -- a third-party dependency that
is not part of /root/reference (the reference does not pin a version; 4.13.0 is installed in the build container).  Its
published algorithm (modules/imgproc/src/resize.cpp, `resizeGeneric_` with `HResizeLinear` / `VResizeLinear`):
  dsize = (round_half_even(W * fx), round_half_even(H * fy));  scale = 1 / fx  (double)
  per destination column dx: f = float((dx + 0.5) * scale - 0.5); sx = floor(f); a = f - sx;
      sx < 0 -> (sx, a) = (0, 0);  sx >= W - 1 -> (sx, a) = (W - 1, 0)          (rows alike)
  horizontal pass in fp32: r[y][dx] = S[y][sx] * (1 - a) + S[y][sx + 1] * a
  vertical pass in fp32:   D[dy][dx] = r[sy][dx] * (1 - b) + r[sy + 1][dx] * b
Parity status: PINNED -- `tests/golden/frames_reference.npz` holds outputs of the reference's own blob.py + cv2 generated
by `tests/golden/make_golden_frames.py` in the build container; `tests/test_frames_cpu.py` checks this restatement against
them (fp32 rounding of the two passes differs from OpenCV's SIMD code by a few ulp, bar 1e-5 of max |x|).
"""
import numpy as np

PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])          # config.py:257 (BGR)


def im_scale_for(h, w, target_size, max_size, cap):
    """minibatch.py / blob.py:43-44 (cap=False: the cap is commented out there) and demo.py:271-274 (cap=True)."""
    size_min, size_max = min(h, w), max(h, w)
    im_scale = float(target_size) / float(size_min)
    if cap and np.round(im_scale * size_max) > max_size:
        im_scale = float(max_size) / float(size_max)
    return im_scale


def resized_shape(h, w, im_scale):
    rnd = lambda v: int(np.rint(v))                                   # cvRound: round half to even
    return rnd(h * im_scale), rnd(w * im_scale)


def _axis_table(n_dst, n_src, scale):
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    a = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo], a[lo] = 0, 0.0
    hi = s >= n_src - 1
    s[hi], a[hi] = n_src - 1, 0.0
    return s, a


def resize_linear_f32(im, im_scale):
    """cv2.resize(im, None, None, fx=im_scale, fy=im_scale, interpolation=cv2.INTER_LINEAR) for float32 HxWxC."""
    im = np.ascontiguousarray(im, dtype=np.float32)
    h, w = im.shape[:2]
    dh, dw = resized_shape(h, w, im_scale)
    scale = 1.0 / im_scale
    sx, ax = _axis_table(dw, w, scale)
    sy, ay = _axis_table(dh, h, scale)
    sx1, sy1 = np.minimum(sx + 1, w - 1), np.minimum(sy + 1, h - 1)
    ax_, ay_ = ax[None, :, None], ay[:, None, None]
    one = np.float32(1.0)

    def hpass(rows):
        return (rows[:, sx] * (one - ax_) + rows[:, sx1] * ax_).astype(np.float32)

    r0, r1 = hpass(im[sy]), hpass(im[sy1])
    return (r0 * (one - ay_) + r1 * ay_).astype(np.float32)


def prep_im_for_blob(im, pixel_means, target_size, max_size, cap=False):
    """blob.py:35-52; `im` uint8 HxWx3 (BGR).  Returns (float32 image, im_scale)."""
    im = im.astype(np.float32)                                        # (copy: the reference mutates a fresh imread)
    im -= pixel_means                                                 # float64 subtract, rounded back to float32
    im_scale = im_scale_for(im.shape[0], im.shape[1], target_size, max_size, cap)
    return resize_linear_f32(im, im_scale), im_scale


def im_list_to_blob(ims):
    """blob.py:20-33."""
    max_shape = np.array([im.shape for im in ims]).max(axis=0)
    blob = np.zeros((len(ims), max_shape[0], max_shape[1], 3), dtype=np.float32)
    for i, im in enumerate(ims):
        blob[i, :im.shape[0], :im.shape[1], :] = im
    return blob


def frames_to_blob(frames, target_size=600, max_size=1000, cap=False, flipped=False, pixel_means=PIXEL_MEANS):
    """_get_image_blob + the loader's NCHW permute: uint8 [n, H, W, 3] -> (data [n, 3, h, w], im_info [n, 3])."""
    ims, scales = [], []
    for im in frames:
        if flipped:
            im = im[:, ::-1, :]                                       # minibatch.py:77-78
        out, s = prep_im_for_blob(im, pixel_means, target_size, max_size, cap)
        ims.append(out)
        scales.append(s)
    blob = im_list_to_blob(ims)
    info = np.array([[blob.shape[1], blob.shape[2], s] for s in scales], dtype=np.float32)
    return np.ascontiguousarray(blob.transpose(0, 3, 1, 2)), info
