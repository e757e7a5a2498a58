"""The reference's checkpoint files, read and written unchanged (checkpoint compatibility).

File name and contents as `trainval_net.py:417-437` writes them and `trainval_net.py:296-308` / `test_net.py:149-165` read
them: ``rfcn_detect_track_{session}_{epoch}_{step}.pth`` = ``torch.save`` of
``{'session', 'epoch' (already + 1), 'model': state_dict, 'optimizer': state_dict, 'pooling_mode', 'class_agnostic'}``;
a file written under ``nn.DataParallel`` holds ``RFCN.module.state_dict()`` -- the same keys.  The module's parameter names,
shapes and order are pinned against the reference's own module in tests/test_checkpoint_cpu.py, so a checkpoint trained
with the reference loads here with ``strict=True`` and the other way round.

The engines compute with packed copies of the weights (BatchNorm folded, fp16 hi/lo split): `load_checkpoint` re-packs a
`D2TTrainEngine` passed as ``engine`` and refuses a frozen `D2TEngine`, which must be built AFTER the load.
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


def load_checkpoint(filename, net, optimizer=None, engine=None, map_location="cpu"):
    """trainval_net.py:296-308 (resume: model + optimizer + cfg.POOLING_MODE) and test_net.py:161-165 (model only).
    Returns the file's dict without the two state_dicts (session, epoch, pooling_mode, class_agnostic)."""
    from .engine import D2TEngine
    if isinstance(engine, D2TEngine) and not hasattr(engine, "refresh_weights"):
        raise ValueError("load_checkpoint: a D2TEngine keeps packed copies of the weights it was built from; load the "
                         "checkpoint into the module first and build the engine afterwards")
    checkpoint = torch.load(filename, map_location=map_location)
    getattr(net, "module", net).load_state_dict(checkpoint['model'])
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint['optimizer'])
    if 'pooling_mode' in checkpoint.keys():
        cfg.POOLING_MODE = checkpoint['pooling_mode']
    if engine is not None:
        engine.refresh_weights()
    return {k: v for k, v in checkpoint.items() if k not in ('model', 'optimizer')}
