"""The one helper of lib/model/utils/net_utils.py that sits on the training step."""
import torch


def _smooth_l1_loss(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma=1.0, dim=[1]):
    """net_utils.py:73-87."""
    sigma_2 = sigma ** 2
    in_box_diff = bbox_inside_weights * (bbox_pred - bbox_targets)
    abs_diff = torch.abs(in_box_diff)
    sign = (abs_diff < 1. / sigma_2).detach().float()
    in_loss = torch.pow(in_box_diff, 2) * (sigma_2 / 2.) * sign + (abs_diff - (0.5 / sigma_2)) * (1. - sign)
    loss = bbox_outside_weights * in_loss
    for i in sorted(dim, reverse=True):
        loss = loss.sum(i)
    return loss.mean()


_CONST_CACHE = {}


def device_const(values, like):
    """a small constant tensor on `like`'s device / dtype, uploaded once (a host->device copy per call would be a
    synchronising operation inside the training step -- and is illegal while a CUDA graph is being captured)"""
    key = (tuple(float(v) for v in values), like.device, like.dtype)
    t = _CONST_CACHE.get(key)
    if t is None:
        t = _CONST_CACHE[key] = torch.tensor(list(key[0]), device=like.device, dtype=like.dtype)
    return t


def save_checkpoint(state, filename):
    """net_utils.py:70-71."""
    torch.save(state, filename)


def adjust_learning_rate(optimizer, decay=0.1):
    """net_utils.py:63-67."""
    for param_group in optimizer.param_groups:
        param_group['lr'] = decay * param_group['lr']
