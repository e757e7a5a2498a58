"""roi_data_layer/minibatch.py on the device against the blobs of the reference's own get_minibatch
(tests/golden/minibatch_reference.npz, made by make_golden_minibatch.py): gt boxes / im_info / img_id exact, image blob
within 1e-6 of max |x| of OpenCV's own code (north-star fp tolerance 1e-4) and bit-identical to the numpy oracle."""
import numpy as np
import pytest
import torch

import common
from oracle import frames as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["landscape", "flipped", "portrait"])
def test_get_minibatch_matches_reference(name, monkeypatch):
    from model.utils.config import cfg
    from roi_data_layer.minibatch import get_minibatch
    g = np.load(common.GOLDEN + "/minibatch_reference.npz")
    monkeypatch.setattr(cfg, "TRAIN_SCALES", (60,))
    im = g[name + "_im"]
    entry = {"image": im, "flipped": bool(g[name + "_flipped"]), "boxes": g[name + "_boxes"],
             "gt_classes": g[name + "_classes"], "track_id": g[name + "_track_id"], "img_id": int(g[name + "_img_id"])}
    for image in (im, torch.from_numpy(im).cuda()):
        blobs = get_minibatch([dict(entry, image=image)], 31)
        want = g[name + "_data"]
        data = blobs["data"].cpu().numpy()
        assert data.shape == want.shape
        assert np.abs(data - want).max() <= 1e-6 * np.abs(want).max()
        mine, _ = oracle.prep_im_for_blob(im[:, ::-1] if entry["flipped"] else im, oracle.PIXEL_MEANS, 60, 1000)
        np.testing.assert_array_equal(data[0], mine)
        np.testing.assert_array_equal(blobs["gt_boxes"], g[name + "_gt_boxes"])
        np.testing.assert_array_equal(blobs["im_info"], g[name + "_im_info"])
        assert blobs["img_id"] == int(g[name + "_img_id"]) and blobs["gt_boxes"].shape == (3, 6)


def test_get_minibatch_reads_files_and_grey_images(tmp_path, monkeypatch):
    cv2 = pytest.importorskip("cv2")
    from model.utils.config import cfg
    from roi_data_layer.minibatch import get_minibatch
    monkeypatch.setattr(cfg, "TRAIN_SCALES", (60,))
    im = common.make_frame(45, 80, 90)
    path = str(tmp_path / "frame.png")
    cv2.imwrite(path, im)
    entry = {"flipped": False, "boxes": np.array([[1, 2, 30, 40]], np.uint16), "gt_classes": np.array([4], np.int32),
             "track_id": np.array([0]), "img_id": 7}
    a = get_minibatch([dict(entry, image=path)], 31)
    b = get_minibatch([dict(entry, image=im)], 31)
    assert torch.equal(a["data"], b["data"])
    grey = get_minibatch([dict(entry, image=im[:, :, 0].copy())], 31)["data"].cpu().numpy()
    want, _ = oracle.prep_im_for_blob(np.repeat(im[:, :, :1], 3, axis=2), oracle.PIXEL_MEANS, 60, 1000)
    np.testing.assert_array_equal(grey[0], want)
