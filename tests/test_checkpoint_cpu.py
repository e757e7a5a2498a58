"""Checkpoint compatibility: the module's state_dict against the reference's own module
(tests/golden/state_dict_reference.json, made by make_golden_state_dict.py from /root/reference), and the reference's
checkpoint file format written / read unchanged (trainval_net.py:296-308, 417-437; test_net.py:149-165)."""
import json

import pytest
import torch

import common  # noqa: F401
from model.faster_rcnn.resnet import resnet
from model.utils.config import cfg


@pytest.fixture(scope="module")
def golden():
    return json.load(open(common.GOLDEN + "/state_dict_reference.json"))


@pytest.mark.parametrize("agnostic", [True, False])
def test_state_dict_equals_the_references_entry_for_entry(golden, agnostic):
    net = resnet(tuple(range(31)), 101, class_agnostic=agnostic).create_architecture()
    mine = [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]
    want = golden["class_agnostic" if agnostic else "per_class"]
    assert len(mine) == len(want) == 640
    if not agnostic:
        # the reference hard-codes the tracking head's 1051 input channels (resnet.py:311), which only fits the
        # class-agnostic model it ships (its per-class module cannot run forward); the mirror sizes it from n_reg_classes
        i = [k for k, _, _ in want].index("corr_bbox_net.weight")
        assert want[i][1] == [6076, 1051, 1, 1] and mine[i][1] == [6076, 2 * 4 * 31 * 49 + 81 + 289 + 289, 1, 1]
        mine[i] = want[i]
    assert mine == want                                   # names, shapes, dtypes and ORDER (optimizer states index by order)


def _reference_style_file(path, golden, seed):
    """what the reference's trainval_net.py would have written: its key list with arbitrary values, SGD state included"""
    g = torch.Generator().manual_seed(seed)
    model = {}
    for k, shape, dtype in golden["class_agnostic"]:
        if dtype == "torch.int64":
            model[k] = torch.tensor(7)
        elif k.endswith("running_var"):
            model[k] = torch.rand(shape, generator=g) + 0.5
        else:
            model[k] = torch.randn(shape, generator=g) * 0.05
    for k in ("weight", "bias"):                          # one module registered twice (resnet.py:297-298): same tensor
        model["RFCN_net." + k] = model["RFCN_base.RFCN_net." + k]
    torch.save({'session': 3, 'epoch': 6, 'model': model, 'optimizer': None, 'pooling_mode': 'align', 'class_agnostic': True}, path)
    return model


def test_reference_checkpoint_loads_strictly_and_round_trips(golden, tmp_path):
    from d2t_b200 import checkpoint
    path = checkpoint.checkpoint_name(str(tmp_path), 3, 5, 9999)
    assert path.endswith("rfcn_detect_track_3_5_9999.pth")
    model = _reference_style_file(path, golden, 1)
    net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture()
    old_mode = cfg.POOLING_MODE
    try:
        meta = checkpoint.load_checkpoint(path, net)
        assert meta == {'session': 3, 'epoch': 6, 'pooling_mode': 'align', 'class_agnostic': True}
        assert cfg.POOLING_MODE == 'align'                # trainval_net.py:306-307
    finally:
        cfg.POOLING_MODE = old_mode
    for k, v in net.state_dict().items():
        assert torch.equal(v, model[k]), k
    # write it back the reference's way (DataParallel-wrapped module, trainval_net.py:417-426) with a live optimizer
    params = [{'params': [p], 'lr': 0.002 if 'bias' in n else 0.001, 'weight_decay': 0.0 if 'bias' in n else 1e-4}
              for n, p in net.named_parameters() if p.requires_grad]        # trainval_net.py:281-287
    opt = torch.optim.SGD(params, momentum=0.9)
    for grp in opt.param_groups[:5]:
        grp['params'][0].grad = torch.ones_like(grp['params'][0])
    opt.step()
    wrapped = torch.nn.Module()
    wrapped.module = net
    out = checkpoint.save_checkpoint(checkpoint.checkpoint_name(str(tmp_path), 3, 6, 10), wrapped, opt, session=3, epoch=6,
                                     class_agnostic=True)
    blob = torch.load(out)
    assert sorted(blob) == ['class_agnostic', 'epoch', 'model', 'optimizer', 'pooling_mode', 'session']
    assert blob['epoch'] == 7 and blob['pooling_mode'] == cfg.POOLING_MODE
    assert list(blob['model']) == [k for k, _, _ in golden["class_agnostic"]]
    net2 = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture()
    params2 = [{'params': [p]} for n, p in net2.named_parameters() if p.requires_grad]
    opt2 = torch.optim.SGD(params2, lr=1.0, momentum=0.0)
    checkpoint.load_checkpoint(out, net2, opt2)
    for (k, a), b in zip(net.state_dict().items(), net2.state_dict().values()):
        assert torch.equal(a, b), k
    assert opt2.param_groups[0]['lr'] == opt.param_groups[0]['lr'] and opt2.param_groups[0]['momentum'] == 0.9
    st, st2 = opt.state_dict()['state'], opt2.state_dict()['state']
    assert st.keys() == st2.keys() and all(torch.equal(st[i]['momentum_buffer'], st2[i]['momentum_buffer']) for i in st)


def test_pretrained_files_load_as_in_the_reference(golden, tmp_path, monkeypatch):
    """resnet.py:259-264 (backbone file: only the trunk's keys are taken) and :304-309 (R-FCN detector checkpoint)."""
    import model.faster_rcnn.resnet as R
    trunk = R.resnet101()
    sd = {k: torch.full_like(v, 0.25) for k, v in trunk.state_dict().items()}
    sd["fc.weight"], sd["fc.bias"] = torch.zeros(1000, 2048), torch.zeros(1000)      # torchvision-style extras: ignored
    torch.save(sd, str(tmp_path / "res101.pth"))
    rfcn_file = str(tmp_path / "rfcn_detect.pth")
    model = _reference_style_file(rfcn_file, golden, 2)
    blob = torch.load(rfcn_file)
    for k in [k for k in blob['model'] if k.startswith("corr_bbox_net")]:           # a detector (no tracking head) file
        del blob['model'][k]
    blob['model']['RCNN_top.unused'] = torch.zeros(3)
    torch.save(blob, rfcn_file)

    net = R.resnet(tuple(range(31)), 101, pretrained=True, class_agnostic=True)
    net.model_path = str(tmp_path / "res101.pth")
    net.create_architecture()
    assert float(net.RFCN_base[6][5].conv2.weight.min()) == 0.25 == float(net.RFCN_base[0].weight.max())
    net = R.resnet(tuple(range(31)), 101, pretrained_rfcn=True, class_agnostic=True)
    net.model_rfcn_path = rfcn_file
    net.create_architecture()
    sd = net.state_dict()
    for k in ("RFCN_base.6.5.conv2.weight", "RFCN_rpn.RPN_Conv.weight", "RFCN_cls_net.bias", "RFCN_net.weight"):
        assert torch.equal(sd[k], model[k]), k
    assert float(sd["corr_bbox_net.weight"].abs().max()) < 0.1                       # freshly initialised (resnet.py:311-312)


def test_net_utils_helpers(tmp_path):
    from model.utils.net_utils import adjust_learning_rate, save_checkpoint
    p = torch.nn.Parameter(torch.zeros(2))
    opt = torch.optim.SGD([{'params': [p], 'lr': 0.01}], momentum=0.9)
    adjust_learning_rate(opt, 0.1)                                                   # net_utils.py:63-67
    assert abs(opt.param_groups[0]['lr'] - 0.001) < 1e-12
    save_checkpoint({'a': 1}, str(tmp_path / "x.pth"))
    assert torch.load(str(tmp_path / "x.pth")) == {'a': 1}


def test_build_optimizer_has_the_references_group_layout(golden):
    """trainval_net.py:280-294: one group per trainable parameter in named_parameters() order (the 'optimizer' entry of a
    reference checkpoint indexes parameters by that position), lr / weight decay per group under cfgs/res101.yml."""
    from d2t_b200.checkpoint import build_optimizer
    net = resnet(tuple(range(31)), 101, class_agnostic=True).create_architecture()
    opt, lr = build_optimizer(net, 0.001)
    names = {id(p): n for n, p in net.named_parameters()}
    mine = [[names[id(g['params'][0])], g['lr'], g['weight_decay']] for g in opt.param_groups]
    want = golden["optimizer_groups_res101_yml"]
    assert len(mine) == len(want) == 107 and lr == 0.001
    assert mine == want
    assert all(len(g['params']) == 1 and g['momentum'] == 0.9 for g in opt.param_groups)
    # a reference optimizer state (same layout) loads, and positions line up with the same parameters
    ref_like = torch.optim.SGD([{'params': [torch.nn.Parameter(torch.zeros_like(g['params'][0]))], 'lr': 0.0005,
                                 'weight_decay': g['weight_decay']} for g in opt.param_groups], momentum=0.9)
    for g in ref_like.param_groups:
        g['params'][0].grad = torch.ones_like(g['params'][0])
    ref_like.step()
    opt.load_state_dict(ref_like.state_dict())
    assert opt.param_groups[0]['lr'] == 0.0005
    assert all(opt.state[g['params'][0]]['momentum_buffer'].shape == g['params'][0].shape for g in opt.param_groups)
    adam, lr_adam = build_optimizer(net, 0.001, "adam")
    assert isinstance(adam, torch.optim.Adam) and abs(lr_adam - 0.0001) < 1e-12 and len(adam.param_groups) == 107
    with pytest.raises(ValueError):
        build_optimizer(net, 0.001, "lbfgs")
