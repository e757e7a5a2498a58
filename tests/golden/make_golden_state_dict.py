"""Generates tests/golden/state_dict_reference.json: every key, shape and dtype of the state_dict of THE REFERENCE'S OWN
`resnet(classes, 101, class_agnostic=True)` (lib/model/faster_rcnn/resnet.py:248-312, rfcn.py:22-64) -- the 'model' entry of
its `rfcn_detect_track_{session}_{epoch}_{step}.pth` checkpoints (trainval_net.py:417-437) -- built on CPU in this container
with the compiled operator extensions shimmed out (they hold no parameters).  /root/reference does not exist on the GPU box.

    python tests/golden/make_golden_state_dict.py
"""
import json
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/lib")


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class _NoParams(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()


def shim(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


shim("easydict", EasyDict=EasyDict)
if "torchvision" not in sys.modules:
    try:
        import torchvision.models  # noqa: F401
    except Exception:   # noqa: BLE001
        shim("torchvision"), shim("torchvision.models")
        sys.modules["torchvision"].models = sys.modules["torchvision.models"]
shim("model.correlation.modules.correlation", Correlation=_NoParams)
shim("model.psroi_pooling.modules.psroi_pool", _PSRoIPooling=_NoParams, PSRoIPool=_NoParams)
shim("model.roi_pooling.modules.roi_pool", _RoIPooling=_NoParams)
shim("model.roi_crop.modules.roi_crop", _RoICrop=_NoParams)
shim("model.roi_align.modules.roi_align", RoIAlignAvg=_NoParams, RoIAlign=_NoParams, RoIAlignMax=_NoParams)
shim("model.nms.nms_gpu", nms_gpu=None)
shim("model.roi_crop.functions.roi_crop", RoICropFunction=object)
shim("model.roi_crop.functions.gridgen", AffineGridGenFunction=object)
nn.init.kaiming_normal = nn.init.kaiming_normal_          # removed aliases the 2018 code calls
nn.init.normal = nn.init.normal_

# rfcn.py and resnet.py mix tabs and spaces (a TabError under Python 3): import both from their source, tabs expanded
def import_expanded(name, path):
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(open(path).read().expandtabs(8), path, "exec"), mod.__dict__)
    return mod


import model.faster_rcnn  # noqa: E402,F401
import_expanded("model.faster_rcnn.rfcn", "/root/reference/lib/model/faster_rcnn/rfcn.py")
resnet = import_expanded("model.faster_rcnn.resnet", "/root/reference/lib/model/faster_rcnn/resnet.py").resnet

from model.utils.config import cfg  # noqa: E402
cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS = [4, 8, 16, 32], [0.5, 1, 2]      # the imagenet_vid set_cfgs (trainval_net.py:165-169)

out = {}
for agnostic in (True, False):
    net = resnet(tuple(range(31)), 101, pretrained=False, class_agnostic=agnostic)
    net.create_architecture()
    sd = net.state_dict()
    out["class_agnostic" if agnostic else "per_class"] = [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]
    print(agnostic, len(sd), "entries,", sum(v.numel() for v in sd.values()), "values")
json.dump(out, open(os.path.join(HERE, "state_dict_reference.json"), "w"))
