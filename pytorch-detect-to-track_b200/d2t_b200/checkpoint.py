"""The reference's checkpoint files, read and written unchanged (checkpoint compatibility).

File name and contents as `trainval_net.py:417-437` writes them and `trainval_net.py:296-308` / `test_net.py:149-165` read
them: ``rfcn_detect_track_{session}_{epoch}_{step}.pth`` = ``torch.save`` of
``{'session', 'epoch' (already + 1), 'model': state_dict, 'optimizer': state_dict, 'pooling_mode', 'class_agnostic'}``;
a file written under ``nn.DataParallel`` holds ``RFCN.module.state_dict()`` -- the same keys.  The module's parameter names,
shapes and order are pinned against the reference's own module in tests/test_checkpoint_cpu.py, so a checkpoint trained
with the reference loads here with ``strict=True`` and the other way round.

The engines compute with packed copies of the parameters (BatchNorm statistics folded into scale / shift, weights split
into fp16 hi/lo pairs), so -- in the order the reference's scripts already follow (trainval_net.py:296-308 loads before the
first step, test_net.py:161-165 before the first forward) -- load the checkpoint into the module FIRST and build
`D2TEngine` / `D2TTrainEngine` afterwards; `load_checkpoint` refuses an engine that is already built (`engines=`), and
`D2TEngine.check_fresh()` reports a module that changed underneath one.
"""
import os

import torch

from model.utils.config import cfg
from model.utils.net_utils import save_checkpoint as _save


def checkpoint_name(output_dir, session, epoch, step):
    """trainval_net.py:418 / test_net.py:149-150."""
    return os.path.join(output_dir, 'rfcn_detect_track_{}_{}_{}.pth'.format(session, epoch, step))


def save_checkpoint(filename, net, optimizer, session, epoch, class_agnostic, pooling_mode=None):
    """trainval_net.py:417-437; `epoch` is the epoch just finished (the file stores epoch + 1, the epoch to resume at).
    `net` may be wrapped (`.module`), as under the reference's --mGPUs."""
    net = getattr(net, "module", net)
    _save({'session': session, 'epoch': epoch + 1, 'model': net.state_dict(), 'optimizer': optimizer.state_dict(),
           'pooling_mode': cfg.POOLING_MODE if pooling_mode is None else pooling_mode, 'class_agnostic': class_agnostic},
          filename)
    return filename


def load_checkpoint(filename, net, optimizer=None, engines=(), map_location="cpu"):
    """trainval_net.py:296-308 (resume: model + optimizer + cfg.POOLING_MODE) and test_net.py:161-165 (model only).
    Returns the file's dict without the two state_dicts (session, epoch, pooling_mode, class_agnostic).
    `engines`: engines already built on `net` -- an error (their packed operands would be stale; see the module docstring)."""
    if engines:
        raise ValueError("load_checkpoint: an engine keeps packed copies of the parameters and folded BatchNorm statistics it "
                         "was built from; load the checkpoint into the module first and build the engine afterwards")
    checkpoint = torch.load(filename, map_location=map_location)
    getattr(net, "module", net).load_state_dict(checkpoint['model'])
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint['optimizer'])
    if 'pooling_mode' in checkpoint.keys():
        cfg.POOLING_MODE = checkpoint['pooling_mode']
    return {k: v for k, v in checkpoint.items() if k not in ('model', 'optimizer')}


def build_optimizer(net, lr=None, optimizer="sgd"):
    """trainval_net.py:275-294: ONE parameter group per trainable parameter, in `named_parameters()` order -- the layout
    the 'optimizer' entry of a reference checkpoint indexes by position -- biases at lr * (DOUBLE_BIAS + 1) with weight
    decay only if BIAS_DECAY, everything else at lr with WEIGHT_DECAY; SGD with TRAIN.MOMENTUM, or Adam (the script's
    own lr * 0.1 for Adam only changes the value it logs, not the groups: :289-291).
    Returns (optimizer, lr)."""
    lr = cfg.TRAIN.LEARNING_RATE if lr is None else lr
    params = []
    for key, value in dict(getattr(net, "module", net).named_parameters()).items():
        if value.requires_grad:
            if 'bias' in key:
                params += [{'params': [value], 'lr': lr * (cfg.TRAIN.DOUBLE_BIAS + 1),
                            'weight_decay': cfg.TRAIN.BIAS_DECAY and cfg.TRAIN.WEIGHT_DECAY or 0}]
            else:
                params += [{'params': [value], 'lr': lr, 'weight_decay': cfg.TRAIN.WEIGHT_DECAY}]
    if optimizer == "adam":
        return torch.optim.Adam(params), lr * 0.1
    if optimizer == "sgd":
        return torch.optim.SGD(params, momentum=cfg.TRAIN.MOMENTUM), lr
    raise ValueError("optimizer must be 'sgd' or 'adam' (trainval_net.py:289-294)")
