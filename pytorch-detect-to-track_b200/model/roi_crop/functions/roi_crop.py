"""Drop-in for lib/model/roi_crop/functions/roi_crop.py:7-21 (and its duplicate crop_resize.py).
``RoICropFunction()(features, grid_yx)``; gradient w.r.t. the grid is zero, as in the reference."""
from d2t_b200 import ops


class RoICropFunction(object):
    def __call__(self, input1, input2):
        assert input1.get_device() == input2.get_device(), "input1 and input2 must on the same device"
        return ops.roi_crop(input1, input2)

    forward = __call__
