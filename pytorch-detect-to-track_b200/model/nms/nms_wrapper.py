"""Drop-in for lib/model/nms/nms_wrapper.py:11-18."""
from model.nms.nms_gpu import nms_gpu


def nms(dets, thresh, force_cpu=False):
    """Greedy NMS over score-sorted ``dets [N, 5]``; ``force_cpu`` is accepted and ignored, as in
    the reference.  Empty input returns ``[]`` (nms_wrapper.py:13-14)."""
    if dets.shape[0] == 0:
        return []
    return nms_gpu(dets, thresh)
