"""lib/roi_data_layer/roibatchLoader.py with the frames kept on the device: one sample = one frame PAIR of a video,
`(data [2, 3, h, w], im_info [2, 3], gt_boxes [2, MAX_NUM_GT_BOXES, 6], num_boxes [2, 1])`, exactly the tuple
`trainval_net.py:355-363` copies into the network's inputs.

Per frame (roibatchLoader.py:104-231): `get_minibatch` (d2t_frames_prep: flip, cast, mean subtraction, resize on the
device), then -- training -- the crop to the batch's aspect ratio when the roidb marks the pair `need_crop` (a random offset
that keeps the boxes, drawn per frame from numpy's global generator in the reference's order), zero padding up to that
ratio, boxes shifted / clamped, degenerate boxes dropped, the rest padded to MAX_NUM_GT_BOXES rows.  Slicing and padding
are tensor views / copies in HBM; only the few box rows are handled on the host, as in the reference."""
import numpy as np
import torch
import torch.utils.data as data

from model.utils.config import cfg
from roi_data_layer.minibatch import get_minibatch


def _crop_start(lo, hi, trim_size, extent):
    """roibatchLoader.py:129-146 / 161-178: start of a `trim_size` window along one axis that keeps the box span [lo, hi]."""
    box_region = hi - lo + 1
    if lo == 0:
        return 0
    if box_region - trim_size < 0:
        s_min, s_max = max(hi - trim_size, 0), min(lo, extent - trim_size)
        return s_min if s_min == s_max else np.random.choice(range(s_min, s_max))
    s_add = int((box_region - trim_size) / 2)
    return lo if s_add == 0 else np.random.choice(range(lo, lo + s_add))


class roibatchLoader(data.Dataset):
    def __init__(self, roidb, ratio_list, ratio_index, batch_size, num_classes, training=True, normalize=None):
        self._roidb = roidb
        self._num_classes = num_classes
        self.max_num_box = cfg.MAX_NUM_GT_BOXES
        self.training = training
        self.normalize = normalize
        self.batch_size = batch_size
        self.ratio_list = ratio_list
        self.ratio_index = ratio_index
        self.data_size = len(self.ratio_list)
        # one aspect ratio per batch (roibatchLoader.py:38-56): the leftmost when the batch is all portrait, the rightmost
        # when all landscape, 1 when it straddles
        self.ratio_list_batch = torch.zeros(self.data_size)
        for i in range(int(np.ceil(len(ratio_index) / batch_size))):
            left_idx, right_idx = i * batch_size, min((i + 1) * batch_size - 1, self.data_size - 1)
            if ratio_list[right_idx] < 1:
                target_ratio = ratio_list[left_idx]
            elif ratio_list[left_idx] > 1:
                target_ratio = ratio_list[right_idx]
            else:
                target_ratio = 1
            self.ratio_list_batch[left_idx:(right_idx + 1)] = target_ratio

    def __len__(self):
        return len(self._roidb)

    def __getitem__(self, index):
        index_ratio = int(self.ratio_index[index]) if self.training else index
        minibatch_db = self._roidb[index_ratio]
        for entry in minibatch_db[:2]:                                         # roibatchLoader.py:99-102
            assert len(entry['track_id']) == len(np.unique(entry['track_id'])), \
                'Cannot have >1 track with same id in same frame.'
        datas, infos, boxes_out, nums = [], [], [], []
        for entry in minibatch_db:
            blobs = get_minibatch([entry], self._num_classes)
            frame = blobs['data'][0]                                           # [h, w, 3] on the device
            im_info = torch.from_numpy(blobs['im_info'])
            h, w = int(frame.size(0)), int(frame.size(1))
            gt = blobs['gt_boxes']
            if not self.training and gt.shape[0] == 0:
                gt = np.ones((1, 6), dtype=np.float32)
            gt = torch.from_numpy(gt)
            if self.training:
                # (the reference divides a Python int by a 0-dim float32 tensor: float32 arithmetic decides floor / ceil)
                ratio = np.float32(self.ratio_list_batch[index].item())
                wf, hf = np.float32(w), np.float32(h)
                if minibatch_db[0]['need_crop']:
                    if ratio < 1.:                                             # much taller than wide: crop the height
                        trim = min(int(np.floor(wf / ratio)), h)
                        y_s = _crop_start(int(torch.min(gt[:, 1])), int(torch.max(gt[:, 3])), trim, h)
                        frame = frame[y_s:(y_s + trim)]
                        gt[:, 1] -= float(y_s)
                        gt[:, 3] -= float(y_s)
                        gt[:, 1].clamp_(0, trim - 1)
                        gt[:, 3].clamp_(0, trim - 1)
                    else:                                                      # much wider than tall: crop the width
                        trim = min(int(np.ceil(hf * ratio)), w)
                        x_s = _crop_start(int(torch.min(gt[:, 0])), int(torch.max(gt[:, 2])), trim, w)
                        frame = frame[:, x_s:(x_s + trim)]
                        gt[:, 0] -= float(x_s)
                        gt[:, 2] -= float(x_s)
                        gt[:, 0].clamp_(0, trim - 1)
                        gt[:, 2].clamp_(0, trim - 1)
                # pad to the batch's ratio (the sizes come from the UNCROPPED blob, as in the reference :188-209)
                if ratio < 1:
                    padded = torch.zeros(int(np.ceil(wf / ratio)), w, 3, device=frame.device)
                    padded[:h] = frame
                    im_info[0, 0] = padded.size(0)
                elif ratio > 1:
                    padded = torch.zeros(h, int(np.ceil(hf * ratio)), 3, device=frame.device)
                    padded[:, :w] = frame
                    im_info[0, 1] = padded.size(1)
                else:
                    trim = min(h, w)
                    padded = frame[:trim, :trim]
                    gt[:, :4].clamp_(0, trim)
                    im_info[0, 0] = trim
                    im_info[0, 1] = trim
                keep = torch.nonzero(((gt[:, 0] == gt[:, 2]) | (gt[:, 1] == gt[:, 3])) == 0).view(-1)
                gt_pad = torch.zeros(self.max_num_box, gt.size(1))
                if keep.numel() != 0:
                    gt = gt[keep]
                    n = min(gt.size(0), self.max_num_box)
                    gt_pad[:n] = gt[:n]
                else:
                    n = 0
                datas.append(padded.permute(2, 0, 1).contiguous().unsqueeze(0))
            else:
                datas.append(frame.permute(2, 0, 1).contiguous().unsqueeze(0))
                gt_pad = torch.zeros(self.max_num_box, gt.size(1))
                n = min(gt.size(0), self.max_num_box)
                gt_pad[:n] = gt[:n]
            infos.append(im_info)
            boxes_out.append(gt_pad.unsqueeze(0))
            nums.append(torch.tensor([[n]], dtype=torch.long))
        dev = datas[0].device
        return (torch.cat(datas, dim=0), torch.cat(infos, dim=0), torch.cat(boxes_out, dim=0),
                torch.cat(nums, dim=0).to(dev))
