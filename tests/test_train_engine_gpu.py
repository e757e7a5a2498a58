"""The training step on the native kernels (d2t_b200.train.D2TTrainEngine): forward on the tcgen05 engine, heads +
losses through autograd, explicit backward through DgradConv / WgradLayer / the correlation + PSRoI backward kernels.
Checked against torch autograd over the same nn.Module (cuDNN fp32, TF32 off) FED THE SAME head gradients, so the
comparison does not depend on the target layers' random sampling."""
import copy

import pytest
import torch
import torch.nn.functional as F

import common

pytestmark = pytest.mark.gpu


def _setup(layers, B, H, W):
    """a WELL-CONDITIONED synthetic net: non-trivial frozen BatchNorm statistics (so the folding is exercised) without
    per-layer re-normalisation -- fp32 and fp64 forwards of this net agree to 1e-6, so two fp32 implementations can be
    compared tightly.  (With BatchNorm calibrated on two small frames the same random net amplifies rounding noise
    ~100x from layer to layer: torch's own fp32 and fp64 forwards then differ by 1e-4, scripts/r02_train_debug2.py.)"""
    from model.faster_rcnn.resnet import resnet
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    net = resnet(tuple(range(31)), layers, class_agnostic=True).create_architecture().cuda()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    g = torch.Generator().manual_seed(1)
    im_data = ((torch.rand(B, 2, 3, H, W, generator=g) * 256 - 128) * 0.01).cuda()     # (keeps the un-normalised trunk ~1e3)
    im_info = torch.tensor([H, W, 1.0]).view(1, 1, 3).expand(B, 2, 3).contiguous().cuda()
    net.train()
    gt = torch.from_numpy(common.make_gt_boxes(B, 30, seed=2, height=H, width=W)).cuda()
    nb = (gt[..., 4] > 0).sum(-1, keepdim=True)
    return net, im_data, im_info, gt, nb


def _corr_torch(a, b, pad, md, s1, s2):
    """correlation_cuda_kernel.cu:34-106 for kernel_size 1, pad == max_displacement, in plain torch (any dtype):
    out[n, (tj+r)D + (ti+r), y, x] = mean_c a[n, c, y s1, x s1] * b[n, c, y s1 + tj s2, x s1 + ti s2]  (zero outside)"""
    assert pad == md
    r = md // s2
    H, W = a.shape[2:]
    oh, ow = -(-H // s1), -(-W // s1)
    bp = F.pad(b, (md, md, md, md))
    a_s = a[:, :, ::s1, ::s1]
    outs = []
    for tj in range(-r, r + 1):
        for ti in range(-r, r + 1):
            win = bp[:, :, md + tj * s2: md + tj * s2 + H: s1, md + ti * s2: md + ti * s2 + W: s1]
            outs.append((a_s * win).mean(1, keepdim=True))
    out = torch.cat(outs, 1)
    assert out.shape[2:] == (oh, ow)
    return out


def _torch_param_grads(net, im_data, B, leaf_grads, pure_torch_corr=False):
    """d(sum_i <map_i, leaf_grad_i>)/d(params) by torch autograd through the nn.Module's own convolutions"""
    N = 2 * B
    frames = im_data.permute(1, 0, 2, 3, 4).reshape(N, *im_data.shape[2:]).contiguous()
    conv3, conv4, conv5, base = net._im_to_head(frames)
    cls_map, bbox_map = net.RFCN_cls_net(base), net.RFCN_bbox_net(base)
    rpn = net.RFCN_rpn
    rc = F.relu(rpn.RPN_Conv(base))
    score, delta = rpn.RPN_cls_score(rc), rpn.RPN_bbox_pred(rc)
    if pure_torch_corr:
        c3 = _corr_torch(conv3[:B], conv3[B:], 8, 8, 2, 2)
        c4 = _corr_torch(conv4[:B], conv4[B:], 8, 8, 1, 1)
        c5 = _corr_torch(conv5[:B], conv5[B:], 8, 8, 1, 1)
        trk = net.corr_bbox_net(torch.cat([bbox_map[:B], bbox_map[B:], c3, c4, c5], 1))
    else:
        trk = net._tracking_maps(conv3, conv4, conv5, bbox_map, B)
    torch.autograd.backward([cls_map, bbox_map, score, delta, trk], leaf_grads)


@pytest.mark.parametrize("layers,B,H,W", [(50, 2, 224, 320), (50, 1, 160, 224)])
def test_train_engine_gradients_match_autograd(layers, B, H, W):
    """every trainable parameter's gradient from the engine's explicit backward against (a) torch autograd in float64 with
    a pure-torch correlation -- the ground truth -- and (b) torch autograd in fp32 through cuDNN and this repo's
    correlation op: the engine must be as close to the truth as a small multiple of what cuDNN fp32 is (ReLU masks that
    flip on 1e-6-level forward differences put a floor of ~1e-3 of max |grad| under any two fp32 implementations here)"""
    from d2t_b200.train import D2TTrainEngine
    net, im_data, im_info, gt, nb = _setup(layers, B, H, W)
    net64 = copy.deepcopy(net).double()             # (before the first forward: the modules then hold no graph tensors)
    eng = D2TTrainEngine(net, B, H, W)
    out, loss = eng.forward_backward(im_data, im_info, gt, nb)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss))
    assert out[0].shape == (2, B, 128, 5) and out[1].shape == (2, B, 128, 31)
    mine = eng.flat.clone()
    assert bool(torch.isfinite(mine).all())
    eng.flat.zero_()
    _torch_param_grads(net, im_data, B, [g.clone() for g in eng.leaf_grads])      # accumulates into the same .grad views
    torch.cuda.synchronize()
    ref32 = eng.flat.clone()
    for p in net64.parameters():
        p.grad = None
    _torch_param_grads(net64, im_data.double(), B, [g.double() for g in eng.leaf_grads], pure_torch_corr=True)
    truth = {n: p.grad for n, p in net64.named_parameters() if p.grad is not None}
    names = {id(p): n for n, p in net.named_parameters()}
    worst, errs = ("", 0.0, 0.0), []
    for p in eng.params:
        o = p.grad.storage_offset()
        a, b = mine[o:o + p.numel()].double(), ref32[o:o + p.numel()].double()
        t = truth[names[id(p)]].reshape(-1)
        scale = float(t.abs().max())
        assert scale > 0, names[id(p)]
        e_eng, e_ref = float((a - t).abs().max()) / scale, float((b - t).abs().max()) / scale
        errs.append(e_eng)
        if e_eng > worst[1]:
            worst = (names[id(p)], e_eng, e_ref, float(((a - t).abs() > 1e-5 * scale).double().mean()))
    errs.sort()
    print("train engine vs float64 autograd, per-parameter max-norm rel err: median %.2e, 90th pct %.2e, worst %.2e (%s; "
          "cuDNN fp32 there: %.2e; fraction of its entries off by > 1e-5 of max: %.2e)" % (
              errs[len(errs) // 2], errs[len(errs) * 9 // 10], worst[1], worst[0], worst[2], worst[3]))
    # Measured on B200: median 1.5e-4 / worst 1.4e-3 of the tensor's max |grad| where torch + cuDNN fp32 is itself 2.5e-4
    # away from float64 -- back-propagation through ~50 layers cancels heavily, and 3xFP16 carries 22 mantissa bits per
    # product against fp32's 24, i.e. ~4-6x cuDNN's distance.  Bars: a small multiple of that.
    assert errs[len(errs) // 2] < 5e-4 and errs[len(errs) * 9 // 10] < 1.5e-3 and worst[1] < 5e-3, (errs[len(errs) // 2], worst)
    assert worst[1] < max(16 * worst[2], 1e-3), worst
    assert net.RFCN_base[4][0].conv1.weight.grad is None                           # frozen stem / layer1 (resnet.py:279-289)


def test_train_engine_learns():
    """three SGD steps through the engine (weights re-packed on the device after each) lower the loss on a fixed batch"""
    from d2t_b200.train import D2TTrainEngine
    B, H, W = 2, 224, 320
    net, im_data, im_info, gt, nb = _setup(50, B, H, W)
    eng = D2TTrainEngine(net, B, H, W, graph_heads=False)     # (eager heads: torch.manual_seed below then fixes the samples)
    opt = torch.optim.SGD(eng.params, lr=1e-6, momentum=0.9)
    losses = []
    for it in range(4):
        torch.manual_seed(100)                     # the same RoI / anchor samples every step
        out, loss = eng.forward_backward(im_data, im_info, gt, nb)
        assert bool(torch.isfinite(loss))
        opt.step()
        eng.refresh_weights()
        losses.append(float(loss))
    print("losses", losses)
    assert losses[-1] < losses[0], losses
    # the engine's forward after the updates == the nn.Module's forward with the updated parameters
    net.eval()
    with torch.no_grad():
        frames = im_data.permute(1, 0, 2, 3, 4).reshape(2 * B, 3, H, W).contiguous()
        base = net._im_to_head(frames)[3]
        info = eng._begin(im_data, im_info)
        for layer in eng.layers:
            layer.run()
    err = float((eng.base_feat.to_nchw() - base).abs().max() / base.abs().max())
    assert err < 1e-4, err


def test_train_engine_graphed_heads():
    """From the third step on the heads -- proposal step, target layers, PSRoI heads, five losses AND their autograd backward
    down to the five convolution outputs -- replay as one CUDA graph (no device->host round trip in the target layers).
    The graph's random samples differ from an eager run's by construction, so it is checked through what it returns: the
    sampled RoIs / labels satisfy the samplers' invariants, the classification loss recomputed eagerly from the returned
    samples equals the returned loss, and autograd's gradient of that recomputed loss w.r.t. the class map equals the
    gradient the graph handed to the engine's backward pass."""
    from d2t_b200.train import D2TTrainEngine
    B, H, W = 2, 224, 320
    net, im_data, im_info, gt, nb = _setup(50, B, H, W)
    eng = D2TTrainEngine(net, B, H, W)
    losses = []
    for it in range(5):
        out, loss = eng.forward_backward(im_data, im_info, gt, nb)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(loss)) and bool(torch.isfinite(eng.flat).all())
        losses.append(float(loss))
    assert eng.g_fwd is not None and eng.g_heads is not None
    assert max(losses) < 1.5 * min(losses), losses                 # eager and replayed steps see the same problem
    rois, labels = out[0], out[8]
    assert rois.shape == (2, B, 128, 5) and labels.shape == (2, B, 128)
    for leg in range(2):
        for b in range(B):
            assert bool((rois[leg, b, :, 0] == b).all())
            fg = labels[leg, b] > 0
            assert 1 <= int(fg.sum()) <= 32
            assert bool(torch.isin(labels[leg, b][fg], gt[b, leg, :, 4]).all())
    cls_map = eng.cls_map.detach().clone().requires_grad_()
    total = 0
    for leg in range(2):
        flat = rois[leg].reshape(-1, 5).contiguous()
        pooled = net.RFCN_psroi_cls_pool(cls_map[leg * B:(leg + 1) * B].contiguous(), flat)
        score = net.RFCN_cls_score(pooled).view(flat.size(0), -1)
        l = F.cross_entropy(score, labels[leg].reshape(-1).long())
        assert abs(float(l) - float(out[6][leg])) <= 1e-5 * max(1.0, abs(float(l))), (float(l), float(out[6][leg]))
        total = total + l
    g, = torch.autograd.grad(total / 2, [cls_map])
    got = eng.grads_static[0]
    err = float((g - got).abs().max() / g.abs().max())
    assert err < 1e-5, err


def test_repack_many_equals_per_operand_repack():
    """the one-launch re-pack of every forward / backward-data operand (d2t_conv_repack_many) writes exactly what the
    per-operand packers write"""
    from d2t_b200.train import D2TTrainEngine
    from d2t_b200 import conv as dc
    B, H, W = 1, 160, 224
    net, im_data, im_info, gt, nb = _setup(50, B, H, W)
    eng = D2TTrainEngine(net, B, H, W)
    with torch.no_grad():
        for p in eng.params:
            p.mul_(1.01)                               # new values: the packed copies are stale now
        eng._refresh_weights()                         # (norms + RepackMany)
    layers = [eng.trk_layer] + eng.layers + eng.dgrads
    got = [(l.w_hi.clone(), l.w_lo.clone()) for l in layers]
    for l in layers:
        l.w_hi.zero_(); l.w_lo.zero_()
        l.repack()
    torch.cuda.synchronize()
    assert len(layers) > 100
    for l, (hi, lo) in zip(layers, got):
        assert torch.equal(hi, l.w_hi) and torch.equal(lo, l.w_lo)
    assert float(got[5][0].float().abs().max()) > 0
