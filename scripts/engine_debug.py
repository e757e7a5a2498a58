import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from model.faster_rcnn.resnet import resnet
from d2t_b200.engine import D2TEngine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
net = resnet(tuple(range(31)), 50, class_agnostic=True).create_architecture().cuda().eval()
B, H, W = 2, 224, 320
g = torch.Generator().manual_seed(1)
im_data = (torch.rand(B, 2, 3, H, W, generator=g) * 256 - 128).cuda()
im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
eng = D2TEngine(net, B, H, W, passes=3, keep_features=True)
out = eng(im_data, im_info)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
with torch.no_grad():
    frames = im_data.permute(1, 0, 2, 3, 4).reshape(2 * B, 3, H, W)
    conv3, conv4, conv5, base = net._im_to_head(frames)
    bbox = net.RFCN_bbox_net(base)
    print("bbox_map", rel(eng.bbox_map, bbox), "per image", [rel(eng.bbox_map[i], bbox[i]) for i in range(4)])
    trk_ref = net._tracking_maps(conv3, conv4, conv5, bbox, B)
    print("trk map", rel(eng.trk_layer.out_nchw, trk_ref))
    tin = eng.trk_in.to_nchw(1051)
    c3 = net.conv3_corr_layer(conv3[:B].contiguous(), conv3[B:].contiguous())
    c4 = net.conv4_corr_layer(conv4[:B].contiguous(), conv4[B:].contiguous())
    c5 = net.conv5_corr_layer(conv5[:B].contiguous(), conv5[B:].contiguous())
    want = torch.cat([bbox[:B], bbox[B:], c3, c4, c5], 1)
    for name, a, b in (("bbox_t", 0, 196), ("bbox_t1", 196, 392), ("c3", 392, 473), ("c4", 473, 762), ("c5", 762, 1051)):
        print(name, rel(tin[:, a:b], want[:, a:b]))
    ref = net(im_data, im_info)
    same = (out[0] - ref[0]).abs().amax(-1) < 1e-2
    print("same rois", float(same.float().mean()))
    for i, nm in ((1, "cls_prob"), (2, "bbox_pred")):
        d = (out[i] - ref[i]).abs().amax(-1)
        print(nm, "max diff on same rois", float(d[same].max()), "overall", float(d.max()))
    d = (out[3] - ref[3]).abs().amax(-1)
    print("tracking_pred max diff", float(d[same[0].reshape(-1)].max()))
    print("bbox_pred scale", float(ref[2].abs().max()), "tracking scale", float(ref[3].abs().max()))
