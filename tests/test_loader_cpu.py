"""Host logic of the roi_data_layer mirror (no GPU): the per-batch aspect ratio rule (roibatchLoader.py:38-56), the
box-keeping crop window (:129-146 / :161-178), and that the image path fails loudly without a CUDA device."""
import numpy as np
import pytest
import torch

import common  # noqa: F401
from roi_data_layer.roibatchLoader import _crop_start, roibatchLoader


def test_ratio_per_batch_rule():
    ratios = [0.5, 0.6, 0.75, 0.9, 1.1, 1.3, 1.5, 2.0, 2.0]
    ld = roibatchLoader([None] * 9, ratios, list(range(9)), 2, 31, training=True)
    # batches (0.5, 0.6) -> leftmost; (0.75, 0.9) -> leftmost; (1.1, 1.3) -> rightmost; (1.5, 2.0) -> rightmost; (2.0,)
    np.testing.assert_allclose(ld.ratio_list_batch.numpy(), [0.5, 0.5, 0.75, 0.75, 1.3, 1.3, 2.0, 2.0, 2.0], rtol=1e-7)
    ld = roibatchLoader([None] * 4, [0.8, 0.95, 1.05, 1.4], list(range(4)), 4, 31, training=True)
    assert ld.ratio_list_batch.tolist() == [1.0] * 4                    # a batch that straddles 1 is made square
    ld = roibatchLoader([None] * 3, [0.9, 1.2, 1.6], list(range(3)), 2, 31, training=True)
    assert ld.ratio_list_batch.tolist() == [1.0, 1.0, pytest.approx(1.6)]
    assert len(ld) == 3


def test_crop_window_keeps_the_boxes_when_it_can():
    rng = np.random.RandomState(0)
    for _ in range(2000):
        extent = int(rng.randint(40, 400))
        trim = int(rng.randint(10, extent + 1))
        lo = int(rng.randint(0, extent - 1))
        hi = int(rng.randint(lo, extent))
        np.random.seed(int(rng.randint(1 << 30)))
        s = int(_crop_start(lo, hi, trim, extent))
        assert 0 <= s
        if lo == 0:
            assert s == 0
        elif hi - lo + 1 < trim:                                        # the span fits: the window contains it, inside the frame
            assert s <= lo and s + trim >= hi and s + trim <= max(extent, hi)
        else:                                                           # it does not: the window starts inside its first half
            assert lo <= s <= lo + max((hi - lo + 1 - trim) // 2, 0)
    # deterministic corners of the reference's rule
    assert _crop_start(0, 50, 20, 100) == 0
    assert _crop_start(90, 99, 100, 100) == 0                           # s_min == s_max == 0: no draw
    assert _crop_start(10, 30, 20, 100) == 10                           # span 21 > 20, half-excess 0: starts at the span


def test_same_draws_as_the_reference_order():
    """one np.random.choice per cropped frame, nothing else: the global generator advances exactly as in the reference"""
    np.random.seed(3)
    a = _crop_start(20, 40, 60, 200)
    np.random.seed(3)
    b = np.random.choice(range(max(40 - 60, 0), min(20, 200 - 60)))
    assert a == b


def test_image_path_needs_a_cuda_device():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from roi_data_layer.minibatch import get_minibatch
    entry = {"image": common.make_frame(20, 30, 1), "flipped": False, "boxes": np.array([[1, 2, 9, 9]], np.uint16),
             "gt_classes": np.array([4], np.int32), "track_id": np.array([0]), "img_id": 7}
    with pytest.raises((RuntimeError, AssertionError, ValueError)):     # no CPU fallback for the frame preparation
        get_minibatch([entry], 31)
