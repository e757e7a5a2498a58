"""Generates tests/golden/frames_reference.npz by running THE REFERENCE'S OWN `lib/model/utils/blob.py` (with the
container's cv2) on seeded uint8 frames -- this container only: /root/reference does not exist on the GPU box.

    python tests/golden/make_golden_frames.py

OpenCV build matters: the container's cv2 4.13.0 dispatches float32 INTER_LINEAR to the closed-source IPP (ippicv 2022.2),
whose coordinate arithmetic differs from OpenCV's own published code (resize.cpp) by up to 1e-4 of max |x| on a
1000-pixel-wide frame.  The golden outputs are produced with `cv2.ipp.setUseIPP(False)` -- OpenCV's own code, which
the numpy restatement reproduces bit for bit on reductions -- and the IPP outputs of the full-size frames are stored
beside them (`*_vals_ipp`) so that the test can bound the distance to that build as well.

Cases (key prefix): small frames stored whole (input + output), one full-size VID frame (720x1280 -> 562x1000, the eval
loops' capped scale; -> 600x1067 uncapped as minibatch.py does) stored as 4096 sampled output positions.
"""
import importlib.util
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common  # noqa: E402  (adds the repo root and the package to sys.path)
from common import make_frame  # noqa: E402
from oracle import frames as oracle  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_blob", "/root/reference/lib/model/utils/blob.py")
ref_blob = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_blob)

# (name, H, W, target_size, max_size, cap, flipped, seed)
SMALL = [
    ("down_cap", 36, 64, 30, 50, True, False, 1),        # capped by the long side (demo.py:273-274)
    ("down_nocap", 36, 64, 30, 50, False, False, 2),     # minibatch.py: cap commented out
    ("up", 23, 31, 40, 1000, False, False, 3),           # magnification, border clamps on both sides
    ("flip", 30, 44, 25, 1000, False, True, 4),          # minibatch.py:77-78
    ("half", 48, 64, 24, 1000, False, False, 5),         # exact 2x decimation (OpenCV switches to its area code)
    ("tall", 50, 21, 33, 60, True, False, 6),            # portrait, capped
    ("same", 20, 28, 20, 1000, False, False, 7),         # scale 1
]
FULL = [("vid_cap", 720, 1280, 600, 1000, True, 11), ("vid_nocap", 720, 1280, 600, 1000, False, 12)]


def reference_prep(im, target, max_size, cap, flipped):
    if flipped:
        im = im[:, ::-1, :]
    if cap:                                                           # demo.py:261-277 (eval loops)
        im_orig = im.astype(np.float32, copy=True)
        im_orig -= oracle.PIXEL_MEANS
        scale = oracle.im_scale_for(im.shape[0], im.shape[1], target, max_size, True)
        return cv2.resize(im_orig, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR), scale
    return ref_blob.prep_im_for_blob(im.copy(), oracle.PIXEL_MEANS, target, max_size)   # minibatch.py:80-81


def main():
    out, worst, worst_ipp = {}, 0.0, 0.0
    cv2.ipp.setUseIPP(False)
    for name, h, w, target, max_size, cap, flipped, seed in SMALL:
        im = make_frame(h, w, seed)
        ref, scale = reference_prep(im, target, max_size, cap, flipped)
        mine, s2 = oracle.prep_im_for_blob(im[:, ::-1] if flipped else im, oracle.PIXEL_MEANS, target, max_size, cap)
        assert mine.shape == ref.shape and s2 == scale, (name, mine.shape, ref.shape)
        worst = max(worst, float(np.abs(mine - ref).max() / np.abs(ref).max()))
        out[name + "_im"], out[name + "_out"] = im, ref.astype(np.float32)
        out[name + "_args"] = np.array([target, max_size, int(cap), int(flipped), scale], dtype=np.float64)
    # two frames of different size through im_list_to_blob (blob.py:20-33)
    a, _ = reference_prep(out["down_cap_im"], 30, 50, True, False)
    b, _ = reference_prep(out["tall_im"], 33, 60, True, False)
    out["blob_pair"] = ref_blob.im_list_to_blob([a, b])
    for name, h, w, target, max_size, cap, seed in FULL:
        im = make_frame(h, w, seed)
        ref, scale = reference_prep(im, target, max_size, cap, False)
        mine, _ = oracle.prep_im_for_blob(im, oracle.PIXEL_MEANS, target, max_size, cap)
        assert mine.shape == ref.shape, (name, mine.shape, ref.shape)
        worst = max(worst, float(np.abs(mine - ref).max() / np.abs(ref).max()))
        rng = np.random.RandomState(seed + 100)
        ys, xs = rng.randint(0, ref.shape[0], 4096), rng.randint(0, ref.shape[1], 4096)
        ys[:4], xs[:4] = [0, 0, ref.shape[0] - 1, ref.shape[0] - 1], [0, ref.shape[1] - 1, 0, ref.shape[1] - 1]
        out[name + "_yx"] = np.stack([ys, xs]).astype(np.int32)
        out[name + "_vals"] = ref[ys, xs].astype(np.float32)
        cv2.ipp.setUseIPP(True)
        ref_ipp, _ = reference_prep(im, target, max_size, cap, False)
        cv2.ipp.setUseIPP(False)
        out[name + "_vals_ipp"] = ref_ipp[ys, xs].astype(np.float32)
        worst_ipp = max(worst_ipp, float(np.abs(mine - ref_ipp).max() / np.abs(ref_ipp).max()))
        out[name + "_args"] = np.array([h, w, target, max_size, int(cap), seed, scale, ref.shape[0], ref.shape[1]],
                                       dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "frames_reference.npz"), **out)
    print("frames_reference.npz written; numpy restatement vs blob.py + cv2: worst max-abs / max-abs = %.2e "
          "(OpenCV's own code), %.2e (IPP-dispatched build)" % (worst, worst_ipp))


if __name__ == "__main__":
    main()
