"""The reference's pure-CPU path for one frame-pair batch -- TEST / BASELINE INFRASTRUCTURE.

The reference has no CPU implementation of its custom ops (correlation.c:3-33 are stubs; PSRoI,
RoIAlign and NMS are CUDA-only) and its model files do not parse under Python 3,
so "the reference's CPU path" is, as BASELINE.md section 4 defines it: the D&T graph of
lib/model/faster_rcnn/rfcn.py:66-250 evaluated leg by leg with stock torch.nn fp32 layers on the
host, with the custom ops replaced by the plain-C restatements of oracle_cpu.c and the proposal
step by the restated proposal_layer.  Only bench.py (cpu_baseline / --impl reference) and tests/
may call this.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cpu as oracle


@torch.no_grad()
def forward_eval(net, im_data, im_info, cfg_key="TEST"):
    """net: model.faster_rcnn.resnet.resnet on the CPU in eval mode.  Returns the first four
    entries of the reference's 10-tuple (rois, cls_prob, bbox_pred, tracking_pred)."""
    from model.utils.config import cfg
    B, L = im_data.shape[:2]
    im_data = im_data.permute(1, 0, 2, 3, 4).contiguous()
    im_info = im_info.permute(1, 0, 2).contiguous()
    anchors = oracle.generate_anchors(scales=tuple(cfg.ANCHOR_SCALES), ratios=tuple(cfg.ANCHOR_RATIOS)).astype(np.float32)
    pre, post, thr = cfg[cfg_key].RPN_PRE_NMS_TOP_N, cfg[cfg_key].RPN_POST_NMS_TOP_N, cfg[cfg_key].RPN_NMS_THRESH
    D_cls, D_loc = net.n_classes, 4 * net.n_reg_classes
    conv3, conv4, conv5, rfcn_bbox, rois, cls_prob, bbox_pred = [], [], [], [], [], [], []
    for leg in range(L):   # rfcn.py:95: the legs run sequentially
        c3, c4, c5, base = net._im_to_head(im_data[leg])
        conv3.append(c3), conv4.append(c4), conv5.append(c5)
        rfcn_cls = net.RFCN_cls_net(base)
        rfcn_bbox.append(net.RFCN_bbox_net(base))
        rpn = net.RFCN_rpn
        x = F.relu(rpn.RPN_Conv(base))
        prob = rpn.cls_prob_from_score(rpn.RPN_cls_score(x), rpn.nc_score_out)
        deltas = rpn.RPN_bbox_pred(x)
        r = oracle.proposal_layer(prob.numpy(), deltas.numpy(), im_info[leg].numpy(), anchors, pre, post, thr)
        rois.append(torch.from_numpy(r))
        flat = r.reshape(-1, 5)
        pc, _ = oracle.psroi_forward(rfcn_cls.numpy(), flat, 1 / 16., 7, 7, 7, D_cls)
        pl, _ = oracle.psroi_forward(rfcn_bbox[leg].numpy(), flat, 1 / 16., 7, 7, 7, D_loc)
        score = F.avg_pool2d(torch.from_numpy(pc), 7).view(flat.shape[0], -1)
        cls_prob.append(F.softmax(score, dim=1).view(B, post, -1))
        bbox_pred.append(F.avg_pool2d(torch.from_numpy(pl), 7).view(B, post, -1))
    corr = lambda a, b, p: torch.from_numpy(oracle.correlation_forward(a.numpy(), b.numpy(), *p))
    feat = torch.cat([rfcn_bbox[0], rfcn_bbox[1], corr(conv3[0], conv3[1], (8, 1, 8, 2, 2)),
                      corr(conv4[0], conv4[1], (8, 1, 8, 1, 1)), corr(conv5[0], conv5[1], (8, 1, 8, 1, 1))], 1)
    trk = net.corr_bbox_net(feat)
    pt, _ = oracle.psroi_forward(trk.numpy(), rois[0].numpy().reshape(-1, 5), 1 / 16., 7, 7, 7, D_loc)
    tracking_pred = F.avg_pool2d(torch.from_numpy(pt), 7).view(B * post, -1)
    return torch.stack(rois), torch.stack(cls_prob), torch.stack(bbox_pred), tracking_pred
