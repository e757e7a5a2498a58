"""Drop-in for lib/model/psroi_pooling/modules/psroi_pool.py:7-18."""
from torch.nn.modules.module import Module

from ..functions.psroi_pool import PSRoIPoolFunction


class _PSRoIPooling(Module):
    def __init__(self, pooled_height, pooled_width, spatial_scale, group_size, output_dim):
        super(_PSRoIPooling, self).__init__()
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.group_size = int(group_size)
        self.output_dim = int(output_dim)

    def forward(self, features, rois):
        return PSRoIPoolFunction(self.pooled_height, self.pooled_width, self.spatial_scale, self.group_size,
                                 self.output_dim)(features, rois)


PSRoIPool = _PSRoIPooling
