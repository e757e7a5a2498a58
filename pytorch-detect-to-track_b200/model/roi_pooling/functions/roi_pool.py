"""Drop-in for lib/model/roi_pooling/functions/roi_pool.py:6-38 (CUDA path only: the CPU
forward of roi_pooling.c is outside the hot path and there is no CPU fallback here)."""
from d2t_b200 import ops


class RoIPoolFunction(object):
    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.feature_size = None
        self.argmax = None
        self.rois = None

    def __call__(self, features, rois):
        return ops.roi_pool(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale, holder=self)

    forward = __call__
