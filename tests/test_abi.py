"""CPU-only checks of the drop-in boundary: libd2t_b200.so loads without a GPU and exports every
function include/d2t_b200.h declares; the host package mirrors the reference operator API."""
import ctypes
import inspect
import subprocess

import pytest

import common  # noqa: F401
from d2t_b200 import _lib


def test_library_loads_and_exports_every_declared_symbol():
    _lib.build()
    handle = ctypes.CDLL(_lib.SO_PATH)
    declared = _lib.declared_symbols()
    assert len(declared) >= 26
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    for s in declared:
        assert s in _lib._SIGS, "no ctypes signature for %s" % s
    assert b"sm_100a" in ctypes.cast(_lib.lib().d2t_version(), ctypes.c_char_p).value or _lib.lib().d2t_version()


def test_reference_launcher_names_are_exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH]).decode()
    for sym in ("Correlation_forward_cuda_kernel", "Correlation_backward_cuda_kernel", "PSROIPoolForwardLauncher",
                "PSROIPoolBackwardLauncher", "ROIAlignForwardLaucher", "ROIAlignBackwardLaucher",
                "ROIPoolForwardLaucher", "ROIPoolBackwardLaucher", "BilinearSamplerBHWD_updateOutput_cuda_kernel",
                "BilinearSamplerBHWD_updateGradInput_cuda_kernel", "nms_cuda_compute"):
        assert " T %s\n" % sym in out, sym


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.SO_PATH]).decode()
    archs = {l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, archs


def test_operator_api_mirrors_reference_signatures():
    from model.correlation.modules.correlation import Correlation
    from model.correlation.functions.correlation import CorrelationFunction
    from model.psroi_pooling.modules.psroi_pool import _PSRoIPooling
    from model.psroi_pooling.functions.psroi_pool import PSRoIPoolFunction, PSRoIPoolingFunction
    from model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    from model.roi_align.functions.roi_align import RoIAlignFunction
    from model.roi_pooling.modules.roi_pool import _RoIPooling
    from model.roi_pooling.functions.roi_pool import RoIPoolFunction
    from model.roi_crop.modules.roi_crop import _RoICrop
    from model.roi_crop.functions.roi_crop import RoICropFunction
    from model.nms.nms_wrapper import nms

    def params(f):
        return [p for p in inspect.signature(f).parameters if p != "self"]

    # lib/model/correlation/modules/correlation.py:6
    assert params(Correlation.__init__) == ["pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply"]
    assert params(CorrelationFunction.__init__) == params(Correlation.__init__)
    # lib/model/psroi_pooling/modules/psroi_pool.py:8
    assert params(_PSRoIPooling.__init__) == ["pooled_height", "pooled_width", "spatial_scale", "group_size", "output_dim"]
    assert PSRoIPoolingFunction is PSRoIPoolFunction
    assert params(RoIAlignFunction.__init__) == ["aligned_height", "aligned_width", "spatial_scale"]
    assert params(RoIPoolFunction.__init__) == ["pooled_height", "pooled_width", "spatial_scale"]
    assert params(nms) == ["dets", "thresh", "force_cpu"]
    m = Correlation(pad_size=8, kernel_size=1, max_displacement=8, stride1=2, stride2=2)
    assert (m.pad_size, m.stride1, m.corr_multiply) == (8, 2, 1)
    assert isinstance(RoIAlignAvg(7, 7, 1 / 16.), RoIAlign) and isinstance(RoIAlignMax(7, 7, 1 / 16.), RoIAlign)
    _RoIPooling(7, 7, 1 / 16.), _RoICrop(), RoICropFunction()


def test_ops_refuse_cpu_tensors():
    import torch
    from model.psroi_pooling.modules.psroi_pool import _PSRoIPooling
    from model.nms.nms_wrapper import nms
    with pytest.raises(ValueError, match="no CPU path"):
        _PSRoIPooling(7, 7, 1 / 16., 7, 4)(torch.zeros(1, 196, 8, 8), torch.zeros(1, 5))
    with pytest.raises(ValueError, match="no CPU path"):
        nms(torch.zeros(3, 5), 0.5)
    assert nms(torch.zeros(0, 5), 0.5) == []   # nms_wrapper.py:13-14


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libd2t_b200.so")
    with pytest.raises(_lib.D2TLibraryMissing):
        _lib.lib()


def test_model_graph_parameter_names_and_shapes():
    from model.faster_rcnn.resnet import resnet
    net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture()
    sd = net.state_dict()
    for k, shape in {"RFCN_base.0.weight": (64, 3, 7, 7), "RFCN_base.6.22.conv2.weight": (256, 256, 3, 3),
                     "RFCN_base.7.0.downsample.0.weight": (2048, 1024, 1, 1), "RFCN_base.RFCN_net.weight": (512, 2048, 3, 3),
                     "RFCN_net.bias": (512,), "RFCN_rpn.RPN_Conv.weight": (512, 512, 3, 3),
                     "RFCN_rpn.RPN_cls_score.weight": (24, 512, 1, 1), "RFCN_rpn.RPN_bbox_pred.weight": (48, 512, 1, 1),
                     "RFCN_cls_net.weight": (1519, 512, 1, 1), "RFCN_bbox_net.weight": (196, 512, 1, 1),
                     "corr_bbox_net.weight": (196, 1051, 1, 1)}.items():
        assert tuple(sd[k].shape) == shape, k
    trainable = sum(p.numel() for p in net.parameters() if p.requires_grad)
    assert abs(trainable - 55.1e6) < 0.1e6          # ~55.1 M trainable fp32 parameters
    net.train()
    assert not net.RFCN_base[4].training and not net.RFCN_base[6][3].bn2.training and net.RFCN_base[6][3].conv2.training
    # layer4 is dilated, not strided (resnet.py:125)
    assert net.RFCN_base[7][0].conv2.dilation == (2, 2) and net.RFCN_base[7][0].conv1.stride == (1, 1)
    assert net.RFCN_base[5][0].conv1.stride == (2, 2)   # stride on the first 1x1 (resnet.py:72-74)
