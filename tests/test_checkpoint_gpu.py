"""Checkpoint files and the engines: a module restored from a reference-format checkpoint gives the
same engine outputs as the module it was saved from; engines are built after the load."""
import pytest
import torch

import common  # noqa: F401

pytestmark = pytest.mark.gpu


def _net(seed, layers=50):
    from model.faster_rcnn.resnet import resnet
    torch.manual_seed(seed)
    net = resnet(tuple(range(31)), layers, class_agnostic=True).create_architecture().cuda().eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    return net


def test_engines_built_on_a_restored_module(tmp_path):
    from d2t_b200 import checkpoint
    from d2t_b200.engine import D2TEngine
    from d2t_b200.train import D2TTrainEngine
    torch.backends.cudnn.allow_tf32 = False
    H, W = 160, 224
    a, b = _net(0), _net(1)
    opt = torch.optim.SGD([p for p in a.parameters() if p.requires_grad], lr=1e-3, momentum=0.9)
    path = checkpoint.save_checkpoint(checkpoint.checkpoint_name(str(tmp_path), 1, 1, 10), a, opt, 1, 1, True)
    g = torch.Generator().manual_seed(5)
    im = (torch.rand(1, 2, 3, H, W, generator=g) * 256 - 128).cuda()
    info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(1, 2, 3).contiguous().cuda()
    eng_a = D2TEngine(a, 1, H, W)
    eng_b_stale = D2TEngine(b, 1, H, W)
    with pytest.raises(ValueError):
        checkpoint.load_checkpoint(path, b, engines=[eng_b_stale])       # a built engine cannot follow a load
    meta = checkpoint.load_checkpoint(path, b)
    assert meta['epoch'] == 2 and meta['class_agnostic'] is True
    with pytest.raises(RuntimeError):
        eng_b_stale.check_fresh()                                        # ... and says so
    eng_b = D2TEngine(b, 1, H, W)
    with torch.no_grad():
        oa, ob = eng_a(im, info), eng_b(im, info)
    torch.cuda.synchronize()
    for x, y in zip(oa[:4], ob[:4]):
        assert torch.equal(x, y)
    # training engine built on a restored module (the resume order of trainval_net.py:296-308)
    c = _net(2)
    checkpoint.load_checkpoint(path, c)
    c.train()
    teng = D2TTrainEngine(c, 1, H, W, use_graphs=False, graph_heads=False)
    with torch.no_grad():
        frames = im.permute(1, 0, 2, 3, 4).reshape(2, 3, H, W).contiguous()
        a.eval()
        base = a._im_to_head(frames)[3]
        teng._begin(im, info)
        for layer in teng.layers:
            layer.run()
    err = float((teng.base_feat.to_nchw() - base).abs().max() / base.abs().max())
    assert err < 1e-4, err                                               # tolerance: 1e-4 of max |x| (north star)
