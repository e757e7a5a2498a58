"""lib/roi_data_layer/minibatch.py with the image work on the device: `get_minibatch(roidb, num_classes)` returns the
same blobs dict -- 'data' [1, h, w, 3] (float32, mean-subtracted, resized; a CUDA tensor here), 'gt_boxes' [n, 6] =
(x1, y1, x2, y2) * im_scale, class, track id, 'im_info' [[h, w, im_scale]], 'img_id' -- for one roidb entry.

A roidb entry is the reference's dict ('image' path, 'flipped', 'boxes', 'gt_classes', 'gt_overlaps', 'track_id',
'img_id'); 'image' may also be the decoded uint8 BGR array (numpy or CUDA tensor) so that a caller with its own decoder
never touches cv2.  Decoding a file is host I/O (cv2.imread, as the reference); everything after it -- the flip, float
cast, mean subtraction and resize of minibatch.py:77-81 / blob.py:35-52 -- is one d2t_frames_prep launch."""
import numpy as np
import numpy.random as npr
import torch

from d2t_b200 import ops
from model.utils.config import cfg


def _decoded(image):
    if isinstance(image, str):
        import cv2                                        # host I/O only (minibatch.py:68)
        image = cv2.imread(image)
        if image is None:
            raise IOError("cannot read image")
    if isinstance(image, np.ndarray):
        if image.ndim == 2:                               # minibatch.py:71-73: grey -> three equal channels
            image = np.concatenate((image[:, :, np.newaxis],) * 3, axis=2)
        image = torch.from_numpy(np.ascontiguousarray(image)).cuda(non_blocking=True)
    return image.contiguous()


def _get_image_blob(roidb, scale_inds):
    """minibatch.py:58-88 -> (blob [n, h, w, 3] on the device, im_scales); the entries' frames must agree in size after
    scaling when n > 1 (the reference pads to the largest; get_minibatch only ever passes one)."""
    processed, im_scales = [], []
    for entry, ind in zip(roidb, scale_inds):
        im = _decoded(entry['image'])
        h, w = im.shape[:2]
        _, _, s = ops.frames_resized_shape(h, w, cfg.TRAIN_SCALES[ind], cfg.TRAIN_MAX_SIZE, cap=False)   # blob.py:43-47
        processed.append(ops.frames_prep(im.view(1, h, w, 3), s, flipped=bool(entry['flipped']),
                                         pixel_means=cfg.PIXEL_MEANS, nhwc=True)[0])
        im_scales.append(s)
    from model.utils.blob import im_list_to_blob
    return im_list_to_blob(processed), im_scales


def get_minibatch(roidb, num_classes):
    """minibatch.py:20-56."""
    num_images = len(roidb)
    random_scale_inds = npr.randint(0, high=len(cfg.TRAIN_SCALES), size=num_images)
    assert cfg.TRAIN.BATCH_SIZE % num_images == 0, \
        'num_images ({}) must divide BATCH_SIZE ({})'.format(num_images, cfg.TRAIN.BATCH_SIZE)
    im_blob, im_scales = _get_image_blob(roidb, random_scale_inds)
    blobs = {'data': im_blob}
    assert len(im_scales) == 1, "Single batch only"
    assert len(roidb) == 1, "Single batch only"
    gt_inds = np.where(roidb[0]['gt_classes'] != 0)[0]    # TRAIN.USE_ALL_GT (config.py:156): every non-background box
    gt_boxes = np.empty((len(gt_inds), 6), dtype=np.float32)
    gt_boxes[:, 0:4] = roidb[0]['boxes'][gt_inds, :] * im_scales[0]
    gt_boxes[:, 4] = roidb[0]['gt_classes'][gt_inds]
    gt_boxes[:, 5] = roidb[0]['track_id'][gt_inds]
    blobs['gt_boxes'] = gt_boxes
    blobs['im_info'] = np.array([[im_blob.shape[1], im_blob.shape[2], im_scales[0]]], dtype=np.float32)
    blobs['img_id'] = roidb[0]['img_id']
    return blobs
