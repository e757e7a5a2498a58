"""Seeded synthetic inputs of the D&T hot path: RoIs, score-sorted boxes, RPN maps, ground-truth
tracks, plus the BatchNorm calibration used by the training bench.  Shared by bench.py, the tests and the golden
generators.  The generators are numpy only -- no torch, no GPU."""
import numpy as np


def make_rois(n_per_img, n_img, height=600, width=1000, seed=21, lo=32.0, hi=512.0, shuffle=False):
    """centre ~ U(image), w,h log-uniform [lo, hi] px, clipped, fractional coords kept; col 0 = image idx."""
    rng = np.random.RandomState(seed)
    out = []
    for b in range(n_img):
        cx = rng.uniform(0, width, n_per_img)
        cy = rng.uniform(0, height, n_per_img)
        w = np.exp(rng.uniform(np.log(lo), np.log(hi), n_per_img))
        h = np.exp(rng.uniform(np.log(lo), np.log(hi), n_per_img))
        x1 = np.clip(cx - w / 2, 0, width - 1)
        x2 = np.clip(cx + w / 2, 0, width - 1)
        y1 = np.clip(cy - h / 2, 0, height - 1)
        y2 = np.clip(cy + h / 2, 0, height - 1)
        out.append(np.stack([np.full(n_per_img, b, np.float64), x1, y1, x2, y2], 1))
    rois = np.concatenate(out, 0).astype(np.float32)
    if shuffle:
        rois = rois[rng.permutation(len(rois))]
    return rois


def make_dets(n, height=600, width=1000, seed=22):
    """Boxes from make_rois with scores ~ U(0,1), sorted by descending score (stable)."""
    rois = make_rois(n, 1, height, width, seed=seed)
    rng = np.random.RandomState(seed + 1000)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    order = np.argsort(-scores, kind="stable")
    return np.concatenate([rois[order, 1:], scores[order, None]], 1).astype(np.float32)


def make_clustered_dets(n, seed=23, height=600, width=1000):
    """Heavily overlapping boxes (jittered copies of a few seeds) -- long suppression chains."""
    rng = np.random.RandomState(seed)
    k = max(1, n // 50)
    base = make_rois(k, 1, height, width, seed=seed)[:, 1:]
    idx = rng.randint(0, k, n)
    boxes = base[idx] + rng.normal(0, 6.0, (n, 4)).astype(np.float32)
    boxes[:, 2] = np.maximum(boxes[:, 2], boxes[:, 0])
    boxes[:, 3] = np.maximum(boxes[:, 3], boxes[:, 1])
    scores = rng.uniform(0, 1, n).astype(np.float32)
    order = np.argsort(-scores, kind="stable")
    return np.concatenate([boxes[order], scores[order, None]], 1).astype(np.float32)


def randn(shape, seed):
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32)


def make_rpn_inputs(B, H, W, A=12, seed=30, im_h=None, im_w=None):
    rng = np.random.RandomState(seed)
    score = rng.standard_normal((B, 2 * A, H, W)).astype(np.float32)
    # softmax over {bg, fg} pairs like rpn.py:66-68
    s = score.reshape(B, 2, A * H, W)
    e = np.exp(s - s.max(1, keepdims=True))
    prob = (e / e.sum(1, keepdims=True)).reshape(B, 2 * A, H, W).astype(np.float32)
    deltas = (rng.standard_normal((B, 4 * A, H, W)) * 0.3).astype(np.float32)
    im_info = np.tile(np.array([[im_h or H * 16, im_w or W * 16, 1.0]], np.float32), (B, 1))
    return prob, deltas, im_info


def make_gt_boxes(B, K=30, seed=2, height=600, width=1000):
    """[B, 2, K, 6] = (x1, y1, x2, y2, cls, track_id): 1-5 boxes per frame, w,h ~ U[40, 400] (clipped),
    same track ids in both frames with +-8 px jitter, one extra unmatched box in some frames."""
    rng = np.random.RandomState(seed)
    gt = np.zeros((B, 2, K, 6), np.float32)
    for b in range(B):
        n = rng.randint(1, 6)
        w = rng.uniform(40, min(400, width * 0.6), n)
        h = rng.uniform(40, min(400, height * 0.6), n)
        x1 = rng.uniform(0, width - w - 1)
        y1 = rng.uniform(0, height - h - 1)
        cls = rng.randint(1, 31, n)
        ids = rng.permutation(n) + 1
        for leg in range(2):
            j = rng.uniform(-8, 8, (n, 4)) if leg else np.zeros((n, 4))
            box = np.stack([x1, y1, x1 + w, y1 + h], 1) + j
            box[:, [0, 2]] = np.clip(box[:, [0, 2]], 0, width - 1)
            box[:, [1, 3]] = np.clip(box[:, [1, 3]], 0, height - 1)
            gt[b, leg, :n, :4] = box
            gt[b, leg, :n, 4] = cls
            gt[b, leg, :n, 5] = ids
        if b % 2 == 1 and n < K:      # a track that ends: present in frame t only
            gt[b, 0, n] = [10, 10, 90, 120, 7, 99]
    return gt


def calibrate_batchnorm(net, frames, chunk=2, var_floor=0.1):
    """Give a randomly initialised trunk the BatchNorm statistics a trained one has: running_mean / running_var := the
    statistics of `frames` [N, 3, H, W] at every BatchNorm (one forward with momentum 1), so activations stay O(1)
    through ~100 layers instead of growing to 1e7 (Kaiming weights with identity BN).  The reference always starts
    from a pretrained trunk (resnet.py:304-309); there are no checkpoints offline.  BN stays frozen afterwards.
    Channels that are (nearly) constant on the calibration frames would get a folded scale of up to 1/sqrt(eps) = 316,
    which only amplifies every implementation's rounding noise: their variance is floored at `var_floor` x the layer's
    median variance."""
    import torch
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    saved = [(m.momentum, m.training) for m in bns]
    for m in bns:
        m.momentum = 1.0
        m.train()
    with torch.no_grad():
        net._im_to_head(frames[:chunk] if chunk else frames)
        for m in bns:
            if m.num_batches_tracked is not None and int(m.num_batches_tracked) > 0:
                m.running_var.clamp_(min=float(m.running_var.median()) * var_floor)
    for m, (mom, tr) in zip(bns, saved):
        m.momentum = mom
        m.train(tr)
    return net


def make_frame(h, w, seed):
    """A uint8 BGR frame [h, w, 3] as cv2.imread would hand it to the frame preparation: blocky content plus noise
    (edges and flat areas)."""
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, size=(h // 4 + 2, w // 4 + 2, 3)).astype(np.float32)
    im = np.kron(base, np.ones((4, 4, 1), np.float32))[:h, :w]
    return np.clip(im + rng.randint(-20, 21, size=(h, w, 3)), 0, 255).astype(np.uint8)
