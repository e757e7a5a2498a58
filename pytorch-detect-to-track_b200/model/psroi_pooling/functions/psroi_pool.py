"""Drop-in for lib/model/psroi_pooling/functions/psroi_pool.py:6-45.

``PSRoIPoolFunction(ph, pw, scale, group, D)(features, rois)``; after the call the object
carries ``output``, ``mappingchannel``, ``rois`` and ``feature_size`` like the reference does
(psroi_pool.py:28-31).  ``PSRoIPoolingFunction`` is the spelling BASELINE.json uses.
"""
from d2t_b200 import ops


class PSRoIPoolFunction(object):
    def __init__(self, pooled_height, pooled_width, spatial_scale, group_size, output_dim):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.group_size = int(group_size)
        self.output_dim = int(output_dim)
        self.output = None
        self.rois = None
        self.feature_size = None

    @property
    def mappingchannel(self):
        """int32 [R, D, ph, pw] channel map, as the reference keeps it after forward (psroi_pool.py:29)."""
        if self.output is None:
            return None
        return ops.psroi_mapping_channel(self.output.size(0), self.pooled_height, self.pooled_width, self.group_size,
                                         self.output_dim, self.output.device)

    def __call__(self, features, rois):
        return ops.psroi_pool(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale,
                              self.group_size, self.output_dim, holder=self)

    forward = __call__


PSRoIPoolingFunction = PSRoIPoolFunction
