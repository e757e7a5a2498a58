"""Drop-in for lib/model/roi_align/functions/roi_align.py:7-47."""
from d2t_b200 import ops


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.rois = None
        self.feature_size = None

    def __call__(self, features, rois):
        if not features.is_cuda:
            raise NotImplementedError   # roi_align.py:28-29
        self.rois = rois
        self.feature_size = features.size()
        return ops.roi_align(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)

    forward = __call__
