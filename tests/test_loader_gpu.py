"""roi_data_layer/roibatchLoader.py on the device against samples of the reference's own loader
(tests/golden/loader_reference.npz, made by make_golden_loader.py): padding to the batch ratio, the random crop that keeps
the boxes (numpy's global generator, same draws in the same order), box shift / clamp / drop / padding, the eval form.
Boxes, counts and im_info exact; frames within 1e-6 of max |x| (OpenCV's own resize; north-star tolerance 1e-4)."""
import numpy as np
import pytest
import torch

import common

pytestmark = pytest.mark.gpu
CASES = ["pad_landscape", "pad_portrait", "square", "crop_wide", "crop_tall", "eval"]


@pytest.mark.parametrize("name", CASES)
def test_loader_sample_matches_reference(name, monkeypatch):
    from model.utils.config import cfg
    from roi_data_layer.roibatchLoader import roibatchLoader
    g = np.load(common.GOLDEN + "/loader_reference.npz")
    monkeypatch.setattr(cfg, "TRAIN_SCALES", (60,))
    ratio, need_crop, training, seed, img_id = g[name + "_cfg"]
    pair = [{"image": g["%s_im%d" % (name, i)], "flipped": False, "boxes": g["%s_boxes%d" % (name, i)],
             "gt_classes": g["%s_gt_classes%d" % (name, i)], "track_id": g["%s_track_id%d" % (name, i)],
             "img_id": int(img_id) + i, "need_crop": int(need_crop)} for i in range(2)]
    loader = roibatchLoader([pair], [float(ratio)], [0], 1, 31, training=bool(training))
    np.random.seed(int(seed))
    data, im_info, gt, num = loader[0]
    want = g[name + "_data"]
    assert data.is_cuda and tuple(data.shape) == want.shape
    assert float(np.abs(data.cpu().numpy() - want).max()) <= 1e-6 * np.abs(want).max()
    np.testing.assert_array_equal(im_info.numpy(), g[name + "_im_info"])
    np.testing.assert_array_equal(gt.numpy(), g[name + "_gt"])
    np.testing.assert_array_equal(num.cpu().numpy(), g[name + "_num"])
    assert tuple(gt.shape) == (2, cfg.MAX_NUM_GT_BOXES, 6) and tuple(num.shape) == (2, 1)


def test_loader_feeds_the_module():
    """the sample tuple is what trainval_net.py:355-363 hands to the network: [2, 3, h, w] -> im_data [1, 2, 3, h, w]"""
    from roi_data_layer.roibatchLoader import roibatchLoader
    pair = [{"image": common.make_frame(90, 160, 300 + i), "flipped": bool(i), "boxes": np.array([[10, 12, 80, 70]], np.uint16),
             "gt_classes": np.array([7], np.int32), "track_id": np.array([1]), "img_id": i, "need_crop": 0} for i in range(2)]
    data, im_info, gt, num = roibatchLoader([pair], [160 / 90.], [0], 1, 31, training=True)[0]
    assert tuple(data.shape) == (2, 3, 600, 1067) and im_info[0].tolist() == [600.0, 1067.0, np.float32(600 / 90.)]
    assert num.view(-1).tolist() == [1, 1] and torch.isfinite(data).all()
