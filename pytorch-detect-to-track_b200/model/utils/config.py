"""The constants the hot path reads from the reference's global ``cfg``
(lib/model/utils/config.py:11-302 after cfgs/res101.yml and the driver overrides at
trainval_net.py:165-169).  Attribute access only -- the YAML/CLI merge machinery is out of
scope; values can be reassigned before a module is constructed,
exactly as the reference drivers do with cfg_from_list.
"""


class _NS(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return self.__dict__[k]

    def __repr__(self):
        return "cfg(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items()))


cfg = _NS(
    ANCHOR_SCALES=[4, 8, 16, 32],          # trainval_net.py:165
    ANCHOR_RATIOS=[0.5, 1, 2],             # config.py:298
    FEAT_STRIDE=[16],                      # config.py:300
    MAX_NUM_GT_BOXES=30,                   # trainval_net.py:165
    POOLING_SIZE=7,                        # config.py:286
    POOLING_MODE='crop',                   # config.py:283 (carried by the checkpoint files; the R-FCN heads use PSRoI)
    TRAIN_SCALES=(600,), TRAIN_MAX_SIZE=1000, TEST_SCALES=(600,), TEST_MAX_SIZE=1000,   # config.py:63,66,168,171 (frame preparation)
    PIXEL_MEANS=(102.9801, 115.9465, 122.7717),   # config.py:257 (BGR)
    N_CLASSES=31,                          # lib/datasets/imagenet_detect.py:28-36
    TRAIN=_NS(
        RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=8,   # config.py:141-147
        RPN_BATCHSIZE=256, RPN_FG_FRACTION=0.5, RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.3,  # :131-139
        RPN_CLOBBER_POSITIVES=False, RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), RPN_POSITIVE_WEIGHT=-1.0,
        BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0,   # res101.yml:8,10; config.py:79-87
        BBOX_NORMALIZE_TARGETS_PRECOMPUTED=True, BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0),
        BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2), BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),     # config.py:113-119
        TRUNCATED=False,
        LEARNING_RATE=0.001, MOMENTUM=0.9, WEIGHT_DECAY=0.0001, DOUBLE_BIAS=False, BIAS_DECAY=False,   # config.py:22-46 + cfgs/res101.yml:11-13
    ),
    TEST=_NS(
        RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=16,    # config.py:191-198
        NMS=0.3,                                                                               # config.py:175
    ),
    RESNET=_NS(FIXED_BLOCKS=1),            # config.py:222
)
