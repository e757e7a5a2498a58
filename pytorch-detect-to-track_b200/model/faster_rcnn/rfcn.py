"""The Detect-to-Track graph of lib/model/faster_rcnn/rfcn.py:22-287, restated for Python 3.

``_RFCN.forward(im_data [B,2,3,H,W], im_info [B,2,3], gt_boxes [B,2,30,6], num_boxes [B,2,1])``
returns the reference's 10-tuple (rfcn.py:248-250).  Differences from the reference, none of
which change results: the two siamese legs share frozen-BN weights, so the trunk and the heads
run ONCE over a 2B batch instead of twice in a python loop (rfcn.py:95-103); proposals for all
2B images come from one batched launch sequence; the three correlations, PSRoI pooling, the
proposal step and NMS are the sm_100a kernels of libd2t_b200.so behind the reference's own
operator classes.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from model.correlation.modules.correlation import Correlation
from model.psroi_pooling.modules.psroi_pool import _PSRoIPooling
from model.rpn.rpn import _RPN
from model.utils.config import cfg
from model.utils.net_utils import _smooth_l1_loss


class _RFCN(nn.Module):
    """R-FCN with the D&T tracking branch."""

    def __init__(self, classes, class_agnostic):
        super(_RFCN, self).__init__()
        self.classes = classes
        self.n_classes = len(classes)
        self.n_reg_classes = 1 if class_agnostic else len(classes)
        self.class_agnostic = class_agnostic
        self.n_bbox_reg = 4 if class_agnostic else len(classes)
        self.RFCN_loss_cls = 0
        self.RFCN_loss_bbox = 0
        self.RFCN_rpn = _RPN(self.dout_base_model)
        self.RFCN_proposal_target = None            # training-only samplers, built on first use
        self.RFCN_tracking_proposal_target = None
        P = cfg.POOLING_SIZE
        self.RFCN_psroi_cls_pool = _PSRoIPooling(P, P, spatial_scale=1.0 / 16.0, group_size=7, output_dim=self.n_classes)
        self.RFCN_psroi_loc_pool = _PSRoIPooling(P, P, spatial_scale=1.0 / 16.0, group_size=7,
                                                 output_dim=4 * self.n_reg_classes)
        self.RFCN_cls_net = nn.Conv2d(512, self.n_classes * 7 * 7, [1, 1], padding=0, stride=1)
        nn.init.normal_(self.RFCN_cls_net.weight.data, 0.0, 0.01)
        self.RFCN_bbox_net = nn.Conv2d(512, 4 * self.n_reg_classes * 7 * 7, [1, 1], padding=0, stride=1)
        nn.init.normal_(self.RFCN_bbox_net.weight.data, 0.0, 0.01)
        self.conv3_corr_layer = Correlation(pad_size=8, kernel_size=1, max_displacement=8, stride1=2, stride2=2)
        self.conv4_corr_layer = Correlation(pad_size=8, kernel_size=1, max_displacement=8, stride1=1, stride2=1)
        self.conv5_corr_layer = Correlation(pad_size=8, kernel_size=1, max_displacement=8, stride1=1, stride2=1)
        self.RFCN_cls_score = nn.AvgPool2d((7, 7), stride=(7, 7))
        self.RFCN_bbox_pred = nn.AvgPool2d((7, 7), stride=(7, 7))
        self.RFCN_tracking_pred = nn.AvgPool2d((7, 7), stride=(7, 7))

    def forward(self, im_data, im_info, gt_boxes=None, num_boxes=None):
        B, L = im_data.size(0), im_data.size(1)
        assert L == 2, "Detect-to-Track runs on frame pairs"
        # leg-major flattening: rows [0, B) are frame t, rows [B, 2B) frame t+tau
        frames = im_data.permute(1, 0, 2, 3, 4).reshape(L * B, *im_data.shape[2:]).contiguous()
        info = im_info.permute(1, 0, 2).reshape(L * B, 3).contiguous()
        conv3, conv4, conv5, base_feat = self._im_to_head(frames)
        rfcn_cls = self.RFCN_cls_net(base_feat)
        rfcn_bbox = self.RFCN_bbox_net(base_feat)

        if self.training:
            from model.rpn.proposal_target_layer_cascade import train_heads   # training-only glue
            return train_heads(self, B, conv3, conv4, conv5, base_feat, rfcn_cls, rfcn_bbox, info, gt_boxes, num_boxes)

        rois_all, _, _ = self.RFCN_rpn(base_feat, info, None, None)            # [2B, R, 5], col 0 = row in 2B batch
        R = rois_all.size(1)
        flat = rois_all.view(-1, 5)
        pooled_cls = self.RFCN_psroi_cls_pool(rfcn_cls, flat)
        pooled_loc = self.RFCN_psroi_loc_pool(rfcn_bbox, flat)
        cls_score = self.RFCN_cls_score(pooled_cls).view(L * B * R, -1)
        cls_prob = F.softmax(cls_score, dim=1).view(L, B, R, -1)
        bbox_pred = self.RFCN_bbox_pred(pooled_loc).view(L, B, R, -1)

        # the reference numbers rois per leg (col 0 in [0, B)), rfcn.py:108-112
        rois = rois_all.view(L, B, R, 5).clone()
        rois[1, :, :, 0] -= B

        tracking_reg_coords = self._tracking_maps(conv3, conv4, conv5, rfcn_bbox, B)
        tracking_rois = rois[0].contiguous().view(-1, 5)                     # rfcn.py:192
        pooled_trk = self.RFCN_psroi_loc_pool(tracking_reg_coords, tracking_rois)
        tracking_pred = self.RFCN_tracking_pred(pooled_trk).view(B * R, -1)

        zero = im_data.new_zeros(L, 1)
        return (rois, cls_prob, bbox_pred, tracking_pred, zero, zero.clone(), zero.clone(), zero.clone(), [],
                im_data.new_zeros(1))

    def _tracking_maps(self, conv3, conv4, conv5, rfcn_bbox, B):
        """rfcn.py:166-175: [bbox_t, bbox_t+tau, corr3, corr4, corr5] -> 1x1 conv."""
        c3 = self.conv3_corr_layer(conv3[:B].contiguous(), conv3[B:].contiguous())
        c4 = self.conv4_corr_layer(conv4[:B].contiguous(), conv4[B:].contiguous())
        c5 = self.conv5_corr_layer(conv5[:B].contiguous(), conv5[B:].contiguous())
        tracking_feat = torch.cat([rfcn_bbox[:B], rfcn_bbox[B:], c3, c4, c5], dim=1)
        return self.corr_bbox_net(tracking_feat)

    def _init_weights(self):
        if not self.pretrained_rfcn:   # rfcn.py:263-266
            for m in (self.RFCN_rpn.RPN_Conv, self.RFCN_rpn.RPN_cls_score, self.RFCN_rpn.RPN_bbox_pred):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def create_architecture(self):
        self._init_modules()
        self._init_weights()
        return self
