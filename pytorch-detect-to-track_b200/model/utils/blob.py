"""lib/model/utils/blob.py on the device: the same two functions, same arguments, same results, with the image kept in
HBM.  `prep_im_for_blob` takes the uint8 BGR frame cv2.imread returns (numpy, or a CUDA uint8 tensor already uploaded) and
returns a CUDA float32 [h, w, 3] tensor; one launch of d2t_frames_prep (csrc/frames.cu) does the cast, the mean
subtraction and OpenCV's float32 bilinear resize.  There is no CPU path."""
import numpy as np
import torch

from d2t_b200 import ops


def _to_device_u8(im):
    if isinstance(im, np.ndarray):
        if im.dtype != np.uint8:
            raise ValueError("prep_im_for_blob takes the uint8 image cv2.imread returns (got %s); the float32 cast and "
                             "the mean subtraction happen on the device" % im.dtype)
        im = torch.from_numpy(np.ascontiguousarray(im)).cuda(non_blocking=True)
    return im.contiguous()


def prep_im_for_blob(im, pixel_means, target_size, max_size):
    """blob.py:35-52: mean subtract and scale an image for use in a blob -> (im [h, w, 3] float32 CUDA, im_scale).
    As in the reference, `max_size` is accepted and NOT applied (the cap is commented out there, blob.py:45-47)."""
    im = _to_device_u8(im)
    h, w = im.shape[:2]
    _, _, im_scale = ops.frames_resized_shape(h, w, target_size, max_size, cap=False)
    means = np.asarray(pixel_means, dtype=np.float64).reshape(-1)
    out = ops.frames_prep(im.view(1, h, w, 3), im_scale, False, means, nhwc=True)
    return out[0], im_scale


def im_list_to_blob(ims):
    """blob.py:20-33: prepared images [h_i, w_i, 3] -> zero-padded blob [n, max h, max w, 3]."""
    max_h, max_w = max(int(im.shape[0]) for im in ims), max(int(im.shape[1]) for im in ims)
    blob = torch.zeros(len(ims), max_h, max_w, 3, dtype=torch.float32, device=ims[0].device)
    for i, im in enumerate(ims):
        blob[i, :im.shape[0], :im.shape[1], :] = im
    return blob
