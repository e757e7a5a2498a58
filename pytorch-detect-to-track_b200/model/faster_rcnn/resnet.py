"""Dilated ResNet-101 trunk + R-FCN head of lib/model/faster_rcnn/resnet.py:66-173, 247-344,
restated for Python 3 with the reference's module / parameter names, so that a reference
checkpoint's state_dict keys line up (RFCN_base.0.weight, RFCN_base.6.22.conv2.weight,
RFCN_base.RFCN_net.weight, RFCN_rpn.RPN_Conv.weight, corr_bbox_net.weight, ...).

Topology facts that matter for parity: the stride of a stage sits on the FIRST
1x1 conv of its first bottleneck (resnet.py:72-74); the stem pool is MaxPool2d(3, 2, padding 0,
ceil_mode=True) (:120); layer4 keeps stride 1 and dilates its 3x3 convs by 2 (:125, :140-156);
the head is a 3x3 dilation-6 conv 2048->512 with bias (:299-301); every BatchNorm is frozen and
always runs in eval mode (:290-295, :314-330).
"""
import math

import torch
import torch.nn as nn

from model.faster_rcnn.rfcn import _RFCN
from model.utils.config import cfg


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1, dilate_first_conv=False):
        super(Bottleneck, self).__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=dilation, bias=False,
                               dilation=dilation)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        residual = x if self.downsample is None else self.downsample(x)
        out += residual
        return self.relu(out)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000):
        self.inplanes = 64
        super(ResNet, self).__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=0, ceil_mode=True)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=2)   # a trous
        for m in self.modules():   # resnet.py:131-137
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample, dilation=dilation)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation))
        return nn.Sequential(*layers)


def resnet50():
    return ResNet(Bottleneck, [3, 4, 6, 3])


def resnet101():
    return ResNet(Bottleneck, [3, 4, 23, 3])


def resnet152():
    return ResNet(Bottleneck, [3, 8, 36, 3])


class resnet(_RFCN):
    """``resnet(classes, 101, pretrained_rfcn, class_agnostic).create_architecture()``.

    Unlike the reference -- which always builds ResNet-101 whatever ``num_layers`` says
    (resnet.py:259) -- ``num_layers`` in {50, 101, 152} is honoured here; 101 is the default and
    the D&T configuration.  ``pretrained`` / ``pretrained_rfcn`` load ``model_path`` / ``model_rfcn_path`` exactly as the
    reference does (resnet.py:259-264, 304-309): only the keys the module owns are taken, the rest of the file is ignored.
    """

    def __init__(self, classes, num_layers=101, pretrained=False, pretrained_rfcn=False, class_agnostic=False):
        self.model_path = 'data/pretrained_model/res101.pth'                  # resnet.py:248-249
        self.model_rfcn_path = 'data/pretrained_model/rfcn_detect.pth'        # trained on ImageNet VID+DET
        self.dout_base_model = 512
        self.num_layers = num_layers
        self.pretrained = pretrained
        self.pretrained_rfcn = pretrained_rfcn
        self.class_agnostic = class_agnostic
        _RFCN.__init__(self, classes, class_agnostic)

    def _init_modules(self):
        trunk = {50: resnet50, 101: resnet101, 152: resnet152}[self.num_layers]()
        if self.pretrained:                                                     # resnet.py:259-264: backbone weights
            state_dict = torch.load(self.model_path, map_location="cpu")
            trunk.load_state_dict({k: v for k, v in state_dict.items() if k in trunk.state_dict()})
        self.RFCN_base = nn.Sequential(trunk.conv1, trunk.bn1, trunk.relu, trunk.maxpool, trunk.layer1, trunk.layer2,
                                       trunk.layer3, trunk.layer4)
        for idx in (0, 1):
            for p in self.RFCN_base[idx].parameters():
                p.requires_grad = False
        assert 0 <= cfg.RESNET.FIXED_BLOCKS < 4
        for blk, idx in ((3, 6), (2, 5), (1, 4)):
            if cfg.RESNET.FIXED_BLOCKS >= blk:
                for p in self.RFCN_base[idx].parameters():
                    p.requires_grad = False
        for m in self.RFCN_base.modules():
            if isinstance(m, nn.BatchNorm2d):
                for p in m.parameters():
                    p.requires_grad = False
        self.RFCN_net = nn.Conv2d(2048, 512, kernel_size=3, padding=6, stride=1, dilation=6)
        self.RFCN_base.add_module("RFCN_net", self.RFCN_net)
        self.RFCN_base.add_module("resnet", trunk.relu)
        nn.init.kaiming_normal_(self.RFCN_net.weight.data)
        if self.pretrained_rfcn:                                                # resnet.py:304-309: an R-FCN detector checkpoint
            pretrained_rfcn_dict = torch.load(self.model_rfcn_path, map_location="cpu")['model']
            pretrained_rfcn_dict = {k: v for k, v in pretrained_rfcn_dict.items() if k in self.state_dict()}
            self.load_state_dict(pretrained_rfcn_dict, strict=False)           # (2018 torch ignored missing keys here)
        n_track_in = 2 * 4 * self.n_reg_classes * 49 + 81 + 289 + 289   # 1051 when class-agnostic (resnet.py:311)
        self.corr_bbox_net = nn.Conv2d(n_track_in, 4 * self.n_reg_classes * 7 * 7, [1, 1], padding=0, stride=1)
        nn.init.normal_(self.corr_bbox_net.weight.data, 0.0, 0.01)

    def train(self, mode=True):
        nn.Module.train(self, mode)
        if mode:   # resnet.py:314-330: stem + layer1 and every BN stay in eval mode
            self.RFCN_base.eval()
            for idx in (5, 6, 7, 8):
                self.RFCN_base[idx].train()
            for m in self.RFCN_base.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def _im_to_head(self, x):
        b = self.RFCN_base
        conv1 = b[3](b[2](b[1](b[0](x))))
        conv2 = b[4](conv1)
        conv3 = b[5](conv2)
        conv4 = b[6](conv3)
        conv5 = b[7](conv4)
        top_feat = b[9](b[8](conv5))
        return conv3, conv4, conv5, top_feat
