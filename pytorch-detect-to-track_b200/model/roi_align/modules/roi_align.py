"""Drop-in for lib/model/roi_align/modules/roi_align.py:6-42."""
from torch.nn.functional import avg_pool2d, max_pool2d
from torch.nn.modules.module import Module

from ..functions.roi_align import RoIAlignFunction


class RoIAlign(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super(RoIAlign, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, self.spatial_scale)(features, rois)


class RoIAlignAvg(RoIAlign):
    def forward(self, features, rois):   # (a+1)x(a+1) samples, then a 2x2 stride-1 average (:26-29)
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1, self.spatial_scale)(features, rois)
        return avg_pool2d(x, kernel_size=2, stride=1)


class RoIAlignMax(RoIAlign):
    def forward(self, features, rois):   # :39-42
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1, self.spatial_scale)(features, rois)
        return max_pool2d(x, kernel_size=2, stride=1)
