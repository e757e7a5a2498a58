#!/usr/bin/env python
"""bench.py -- Detect-to-Track per-frame-pair hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (host cores)

A "step" is one eval-mode forward of the D&T graph (Res-101 siamese trunk -> RPN proposals + NMS ->
PSRoI cls/loc -> 3 correlations -> tracking PSRoI) over the per-GPU batch of synthetic 600x1000
frame-pairs (BASELINE.json configs[1]: batch = 2 pairs on 1 GPU; N GPUs = N x 2 pairs, sharded by
image, no data-path collective -> weak scaling).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-detect-to-track_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "frame-pairs/sec (Res-101 D&T, 600px)"
H, W = 600, 1000
PAIRS_PER_GPU = 2
CLASSES = tuple(range(31))


WORKLOAD = ("Res-101 D&T eval forward, 600x1000 frame-pairs, 300 RoIs/frame, PSRoI + correlation "
            "(BASELINE.json configs[1])")


def shared_config():
    """`config` of the JSON line: the same literal dict in both arms (this repo's path and --impl reference)"""
    return {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU,
            "parallelism": "frame-pairs sharded over ranks, no data-path collective in the eval forward",
            "l2": "GPU arm: 512 MB buffer rewritten between timed steps (> 126 MB L2); CPU arm: not applicable"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        # tensor peak: the BURST figure -- the step is ~6 ms of short kernels at the maximum SM clock with the board far below
        # its power cap (see "clocks" in the output), not a seconds-long power-capped loop; the sustained figure is reported too
        peaks.sustained = float(d.get("bf16_tflops_sustained", d["bf16_tflops"]))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    peaks.sustained = 1400.0
    return 6650.0, 1590.0, "fallback"


def ncu_traffic():
    """DRAM bytes per launch from the committed ncu captures (profiles/r02_traffic.json, else round 1's); bench.py cannot
    run ncu itself."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            d["_file"] = "profiles/" + name
            return d
        except (OSError, ValueError):
            continue
    return {}


def make_inputs(pairs, seed):
    g = torch.Generator().manual_seed(seed)
    im_data = torch.rand(pairs, 2, 3, H, W, generator=g) * 256.0 - 128.0     # mean-subtracted-image-like
    im_info = torch.tensor([float(H), float(W), 1.0]).view(1, 1, 3).expand(pairs, 2, 3).contiguous()
    return im_data, im_info


def build_net(layers=101):
    from model.faster_rcnn.resnet import resnet
    torch.manual_seed(3)
    return resnet(CLASSES, layers, class_agnostic=True).create_architecture().eval()


class ClockSampler(object):
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def flush_l2(buf):
    buf.add_(1.0)   # 512 MB read+write > 126 MB L2


def time_kernel(fn, iters, flush):
    """Average device time (ms) of fn() over `iters` launches, CUDA events on the current stream,
    L2 flushed between launches, host launch latency hidden behind a device-side spin."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush_l2(flush)
        # a ~50 us device-side spin ahead of the start event: the host enqueues event, launch(es) and end event while the
        # GPU is still busy, so the interval is the kernels' device time and not the host's time to reach the launch
        # (Python + ctypes: 5-20 us, comparable to the kernels timed here)
        torch.cuda._sleep(100000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def op_microbench(flush, hbm_gbs):
    """BASELINE configs[3] and [4] at a fixed batch: PSRoI (the roofline kernel of the metric),
    conv4 correlation, proposal-size NMS.  Algorithmic bytes: SURVEY.md section 8d / DESIGN.md."""
    from d2t_b200 import ops
    from d2t_b200 import synth as common
    from d2t_b200._lib import lib
    out = {}
    B, D, R = 2, 30, 2000
    torch.manual_seed(20)
    feat = torch.randn(B, D * 49, 38, 63, device="cuda")
    rois = torch.from_numpy(common.make_rois(R, B, seed=21)).cuda()
    top = torch.empty(B * R, D, 7, 7, device="cuda")
    ws = torch.empty(lib().d2t_psroi_workspace_bytes(B * R, B, 7, 7), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def psroi():   # the C-ABI call itself: prep kernel + planes kernel, mapping_channel skipped
        lib().d2t_psroi_forward(feat.data_ptr(), B, D * 49, 38, 63, rois.data_ptr(), B * R, 1 / 16., 7, 7, 7, D,
                                top.data_ptr(), None, ws.data_ptr(), ws.numel(), st)
    ms = time_kernel(psroi, 30, flush)
    alg = 4.0 * (D * 49 * 2394 + 5 * R + R * D * 49) * B
    out["psroi_fwd"] = {"shape": "feat[%d,%d,38,63] rois %d/img D=%d" % (B, D * 49, R, D), "ms": ms,
                        "algorithmic_bytes": alg, "gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / hbm_gbs}
    # batch sweep of config 5 (same generator): the integer-table kernel keeps 3 CTAs per SM busy once there are more
    # items (image, class, bin row) than CTA slots
    for Bs in (8, 32):
        f_s = torch.randn(Bs, D * 49, 38, 63, device="cuda")
        r_s = torch.from_numpy(common.make_rois(R, Bs, seed=21)).cuda()
        t_s = torch.empty(Bs * R, D, 7, 7, device="cuda")
        w_s = torch.empty(lib().d2t_psroi_workspace_bytes(Bs * R, Bs, 7, 7), dtype=torch.uint8, device="cuda")
        ms_s = time_kernel(lambda: lib().d2t_psroi_forward(f_s.data_ptr(), Bs, D * 49, 38, 63, r_s.data_ptr(), Bs * R, 1 / 16., 7,
                                                           7, 7, D, t_s.data_ptr(), None, w_s.data_ptr(), w_s.numel(), st), 10, flush)
        alg_s = alg / B * Bs
        out["psroi_fwd_b%d" % Bs] = {"ms": ms_s, "algorithmic_bytes": alg_s, "gbs": alg_s / ms_s / 1e6,
                                     "frac_hbm": alg_s / ms_s / 1e6 / hbm_gbs}
        del f_s, r_s, t_s, w_s
    lib().d2t_psroi_set_mode(0, 0)             # the exactly-rounded fp64-table kernel, for comparison
    ms_x = time_kernel(psroi, 10, flush)
    lib().d2t_psroi_set_mode(-1, 0)
    out["psroi_fwd_fp64_tables"] = {"ms": ms_x, "gbs": alg / ms_x / 1e6, "frac_hbm": alg / ms_x / 1e6 / hbm_gbs}
    gt = torch.randn_like(top)
    grad = torch.empty_like(feat)

    def psroi_b():
        lib().d2t_psroi_backward(gt.data_ptr(), B, D * 49, 38, 63, rois.data_ptr(), B * R, 1 / 16., 7, 7, 7, D,
                                 grad.data_ptr(), 0, ws.data_ptr(), ws.numel(), st)
    ms = time_kernel(psroi_b, 20, flush)
    out["psroi_bwd"] = {"kernel": "psroi_prep + psroi_bwd_amax + psroi_bwd_limb (two-limb integer difference tables)", "ms": ms,
                        "algorithmic_bytes": alg, "gbs": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / hbm_gbs}
    lib().d2t_psroi_set_mode(-1, 2)            # the fp64 difference tables (shared-memory CAS loops) it replaced
    ms_x = time_kernel(psroi_b, 10, flush)
    lib().d2t_psroi_set_mode(-1, 0)
    out["psroi_bwd_fp64_tables"] = {"ms": ms_x, "gbs": alg / ms_x / 1e6, "frac_hbm": alg / ms_x / 1e6 / hbm_gbs}
    for name, (C_, Hh, Ww, p, Bc) in {"corr_conv4": (1024, 38, 63, (8, 1, 8, 1, 1), 2), "corr_conv5": (2048, 38, 63, (8, 1, 8, 1, 1), 2),
                                      "corr_conv3": (512, 75, 125, (8, 1, 8, 2, 2), 2)}.items():
        a, b = torch.randn(Bc, C_, Hh, Ww, device="cuda"), torch.randn(Bc, C_, Hh, Ww, device="cuda")
        oc, oh, ow = ops.correlation_shape(Hh, Ww, *p)
        from d2t_b200 import conv as dc
        # tensor-core kernel on the engine's native layout (split NHWC in, NCHW out) ...
        layer = dc.CorrLayer(dc.ActTensor.from_nchw(a), dc.ActTensor.from_nchw(b), p[0], p[2], p[3], passes=16, want_nchw=True)
        ms = time_kernel(layer.run, 20, flush)
        layer3 = dc.CorrLayer(dc.ActTensor.from_nchw(a), dc.ActTensor.from_nchw(b), p[0], p[2], p[3], passes=3, want_nchw=True)
        ms_tf32 = time_kernel(layer3.run, 10, flush)
        # ... through the reference-layout operator (adds the two NCHW -> split-NHWC re-layouts) ...
        ms_api = time_kernel(lambda: ops.correlation_forward(a, b, *p), 10, flush)
        # ... and the fp32 SIMT kernel it replaced
        o = torch.empty(Bc, oc, oh, ow, device="cuda")
        ms_simt = time_kernel(lambda: lib().d2t_correlation_forward(a.data_ptr(), b.data_ptr(), Bc, C_, Hh, Ww, *p, o.data_ptr(), st), 5, flush)
        touched = (oh * ow) if p[3] > 1 else Hh * Ww      # stride-2 lattice reads 1/4 of the elements
        alg = 4.0 * (2 * C_ * touched + oc * oh * ow) * Bc
        flops = 2.0 * oc * oh * ow * C_ * Bc
        out[name] = {"batch": Bc, "ms": ms, "algorithmic_bytes": alg, "gbs": alg / ms / 1e6,
                     "frac_hbm": alg / ms / 1e6 / hbm_gbs, "tflops_useful": flops / ms / 1e9,
                     "kernel": "conv_igemm CORR mode, 3xFP16 (kind::f16); 3xTF32 variant: %.4f ms" % ms_tf32,
                     "ms_3xtf32": ms_tf32, "ms_operator_api_nchw": ms_api, "ms_fp32_simt_kernel": ms_simt}
        # backward through the reference-layout operator, both inputs: banded GEMMs on the tensor cores (CORRB mode; re-layouts
        # and operand packers included) against the exact-adjoint fp32 SIMT gather kernels it replaced
        go = torch.randn(Bc, oc, oh, ow, device="cuda")
        ms_bwd = time_kernel(lambda: ops.correlation_backward(a, b, go, *p), 5, flush)
        ops.TENSOR_CORE_CORRELATION = False
        ms_bwd_simt = time_kernel(lambda: ops.correlation_backward(a, b, go, *p), 3, flush)
        ops.TENSOR_CORE_CORRELATION = True
        out[name]["bwd_operator_api_ms_both_inputs"] = ms_bwd
        out[name]["bwd_simt_gather_ms_both_inputs"] = ms_bwd_simt
        out[name]["bwd_frac_hbm"] = 4.0 * (4 * C_ * Hh * Ww + oc * oh * ow) * Bc / ms_bwd / 1e6 / hbm_gbs
        del a, b, o, go, layer, layer3
    # fused PSRoI + 7x7 vote (+ softmax) on the model's own head shapes (4 frames x 300 RoIs, D = 31 classes / 4 deltas)
    for name, (Dv, sm) in {"psroi_vote_cls_softmax": (31, True), "psroi_vote_bbox": (4, False)}.items():
        fv = torch.randn(4, Dv * 49, 38, 63, device="cuda")
        rv = torch.from_numpy(common.make_rois(300, 4, seed=23)).cuda()
        ms_v = time_kernel(lambda: ops.psroi_vote(fv, rv, 7, 7, 1 / 16., 7, Dv, softmax=sm), 20, flush)
        ms_u = time_kernel(lambda: ops.psroi_forward(fv, rv, 7, 7, 1 / 16., 7, Dv)[0].mean((2, 3)), 10, flush)
        alg_v = 4.0 * (4 * Dv * 49 * 2394 + 5 * 1200 + 1200 * Dv)
        out[name] = {"shape": "feat[4,%d,38,63] rois 300/img" % (Dv * 49), "ms": ms_v, "ms_unfused_psroi_plus_mean": ms_u,
                     "algorithmic_bytes": alg_v, "frac_hbm": alg_v / ms_v / 1e6 / hbm_gbs}
        del fv, rv
    # the three RoI operators that are not on the D&T graph (roi_align / roi_pooling / roi_crop: SURVEY 8 rows a8-a10), on a
    # conv4-sized map: 2 images x 1024 channels x 38 x 63, 256 RoIs, 7x7 bins
    fr = torch.randn(2, 1024, 38, 63, device="cuda")
    rr = torch.from_numpy(common.make_rois(128, 2, seed=24)).cuda()
    alg_r = 4.0 * (fr.numel() + rr.numel() + 256 * 1024 * 49)
    top_a = ops.roi_align_forward(fr, rr, 7, 7, 1 / 16.)
    ms_f = time_kernel(lambda: ops.roi_align_forward(fr, rr, 7, 7, 1 / 16.), 10, flush)
    ga = torch.randn_like(top_a)
    ms_b = time_kernel(lambda: ops.roi_align_backward(ga, rr, tuple(fr.shape), 7, 7, 1 / 16.), 10, flush)
    ms_d = time_kernel(lambda: ops.roi_align_backward(ga, rr, tuple(fr.shape), 7, 7, 1 / 16., deterministic=True), 10, flush)
    out["roi_align_256rois"] = {"fwd_ms": ms_f, "bwd_ms": ms_b, "bwd_deterministic_ms": ms_d,
                                "fwd_frac_hbm": alg_r / ms_f / 1e6 / hbm_gbs, "bwd_frac_hbm": alg_r / ms_b / 1e6 / hbm_gbs}
    top_p, arg_p = ops.roi_pool_forward(fr, rr, 7, 7, 1 / 16.)
    ms_f = time_kernel(lambda: ops.roi_pool_forward(fr, rr, 7, 7, 1 / 16.), 10, flush)
    ms_b = time_kernel(lambda: ops.roi_pool_backward(ga, arg_p, rr, tuple(fr.shape), 7, 7, 1 / 16.), 10, flush)
    ms_d = time_kernel(lambda: ops.roi_pool_backward(ga, arg_p, rr, tuple(fr.shape), 7, 7, 1 / 16., deterministic=True), 10, flush)
    out["roi_pool_256rois"] = {"fwd_ms": ms_f, "bwd_ms": ms_b, "bwd_deterministic_ms": ms_d,
                               "fwd_frac_hbm": alg_r / ms_f / 1e6 / hbm_gbs, "bwd_frac_hbm": alg_r / ms_b / 1e6 / hbm_gbs}
    gridc = (torch.rand(2, 38, 63, 2, device="cuda") * 2 - 1).contiguous()
    oc_ = ops.roi_crop_forward(fr, gridc)
    ms_f = time_kernel(lambda: ops.roi_crop_forward(fr, gridc), 10, flush)
    gc = torch.randn_like(oc_)
    ms_b = time_kernel(lambda: ops.roi_crop_backward(fr, gridc, gc), 10, flush)
    alg_c = 4.0 * (2 * fr.numel() + gridc.numel())
    ms_d = time_kernel(lambda: ops.roi_crop_backward(fr, gridc, gc, deterministic=True), 10, flush)
    out["roi_crop_38x63_grid"] = {"fwd_ms": ms_f, "bwd_ms": ms_b, "bwd_deterministic_ms": ms_d, "fwd_frac_hbm": alg_c / ms_f / 1e6 / hbm_gbs,
                                  "bwd_frac_hbm": 1.5 * alg_c / ms_b / 1e6 / hbm_gbs}
    del fr, rr, top_a, ga, top_p, arg_p, gridc, oc_, gc
    # detection decode + per-class NMS after the network (test_net.py:232-301), 4 frames x 30 classes
    from d2t_b200 import detect
    g = torch.Generator().manual_seed(50)
    L_, B_, R_, C_ = 2, 2, 300, 31
    xy = torch.rand(L_, B_, R_, 2, generator=g) * torch.tensor([800.0, 480.0])
    wh = 40 + torch.rand(L_, B_, R_, 2, generator=g) * 200
    d_rois = torch.cat([torch.zeros(L_, B_, R_, 1), xy, (xy + wh).clamp(max=599.0)], -1).cuda()
    d_prob = torch.softmax(torch.randn(L_, B_, R_, C_, generator=g) * 3.0, -1).cuda()
    d_pred = torch.randn(L_, B_, R_, 4, generator=g).cuda()
    d_info = torch.tensor([600.0, 1000.0, 1.0]).view(1, 1, 3).expand(B_, L_, 3).contiguous().cuda()
    ms_b = time_kernel(lambda: detect.per_class_detections(d_rois, d_prob, d_pred, d_info, thresh=0.05), 10, flush)
    t0 = time.time()
    detect.detect_reference_loop(d_rois, d_prob, d_pred, d_info, 0.05, 0.3, 0)
    torch.cuda.synchronize()
    out["detect_postproc_4frames_x30classes"] = {"ms_batched_device": ms_b, "ms_per_class_loop_wall": (time.time() - t0) * 1e3,
                                                 "note": "batched = one sort + gather + d2t_nms_batched over the (frame, class) axis; "
                                                         "loop = the reference's class-by-class nms with a host round trip each"}
    # frame preparation (blob.py:20-52 + minibatch.py:77-78 + the loader's permute) for the step's 4 frames
    fr_u8 = torch.from_numpy(np.stack([common.make_frame(720, 1280, 60 + i) for i in range(4)])).cuda()
    fh, fw, fs = ops.frames_resized_shape(720, 1280, 600, 1000, cap=True)
    blob = torch.empty(4, 3, fh, fw, device="cuda")
    ms = time_kernel(lambda: ops.frames_prep(fr_u8, fs, out=blob), 10, flush)
    alg_f = float(fr_u8.numel() + 4 * blob.numel())
    out["frames_prep_4x720x1280_to_%dx%d" % (fh, fw)] = {"ms": ms, "algorithmic_bytes": alg_f, "gbs": alg_f / ms / 1e6,
                                                         "frac_hbm": alg_f / ms / 1e6 / hbm_gbs}
    del fr_u8, blob
    for n in (6000, 12000):
        dets = torch.from_numpy(np.stack([common.make_dets(n, seed=22 + i) for i in range(4)])).cuda()
        ms = time_kernel(lambda: ops.nms_batched(dets, 0.7, max_keep=300 if n == 6000 else 2000), 10, flush)
        out["nms_%d_x4img" % n] = {"ms": ms, "us_per_image": ms * 1e3 / 4}
    return out


def cpu_path_pairs_per_sec(threads, steps, warmup, budget_s=240.0, pairs=1):
    """The reference's CPU path (oracle/cpu_graph.py) on `pairs` 600x1000 frame-pairs per step."""
    from oracle import cpu as oracle
    from oracle import cpu_graph
    torch.set_num_threads(threads)
    oracle.lib().oracle_set_threads(threads)
    net = build_net(101)
    im_data, im_info = make_inputs(pairs, seed=1)
    t0 = time.time()
    cpu_graph.forward_eval(net, im_data, im_info)
    first = time.time() - t0
    w_run = max(0, min(warmup - 1, int(budget_s * 0.25 / max(first, 1e-3))))
    k_run = max(1, min(steps, int(budget_s * 0.75 / max(first, 1e-3))))
    for _ in range(w_run):
        cpu_graph.forward_eval(net, im_data, im_info)
    times = []
    for _ in range(k_run):
        t0 = time.time()
        cpu_graph.forward_eval(net, im_data, im_info)
        times.append(time.time() - t0)
    sec = float(np.mean(times))
    return pairs / sec, sec, k_run, w_run + 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    value, sec, k_run, w_run = cpu_path_pairs_per_sec(threads, args.steps, args.warmup, pairs=PAIRS_PER_GPU)
    sample = "%d frame-pairs (%d frames 600x1000) per step, full eval graph on host cores; %d timed steps" % (
        PAIRS_PER_GPU, 2 * PAIRS_PER_GPU, k_run)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frame-pairs/s", "n_gpus": args.gpus,
            "steps": k_run, "warmup": w_run, "steps_requested": args.steps, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": shared_config(),
            "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def check_finite(outs, what):
    """a timing of NaN/Inf is not a measurement: fail the run (rc != 0) instead of printing a line"""
    bad = [i for i, t in enumerate(outs) if isinstance(t, torch.Tensor) and t.is_floating_point() and not bool(torch.isfinite(t).all())]
    if bad:
        raise SystemExit("bench.py: non-finite values in %s (outputs %s) -- refusing to report a throughput" % (what, bad))


def parity_vs_torch(net, engine, im_dev, info_dev):
    """One forward of the SAME nn.Module by torch (cuDNN fp32, TF32 off) on the bench inputs, outside every timed region,
    against the engine's outputs: the line carries the error of the numbers it timed (tests/test_model_gpu.py holds the
    full check, float64 included)."""
    N = 2 * im_dev.size(0)
    out = engine(im_dev, info_dev)
    frames = im_dev.permute(1, 0, 2, 3, 4).reshape(N, 3, H, W).contiguous()
    with torch.no_grad():
        base = net._im_to_head(frames)[3]
        ref = net(im_dev, info_dev)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    same = (out[0] - ref[0]).abs().amax(-1) < 1e-2
    sel = same.view(-1)
    res = {"comparator": "the same nn.Module run by torch (cuDNN fp32, allow_tf32=False) on the bench inputs",
           "base_feat_max_rel_err": rel(engine.base_feat.to_nchw(), base),
           "identical_proposals_frac": float(same.float().mean()),
           "cls_prob_max_abs_err": float((out[1].reshape(-1, out[1].size(-1))[sel] - ref[1].reshape(-1, ref[1].size(-1))[sel]).abs().max()),
           "bbox_pred_max_rel_err": float((out[2].reshape(-1, 4)[sel] - ref[2].reshape(-1, 4)[sel]).abs().max() / ref[2].abs().max()),
           "tolerance": 1e-4}
    res["ok"] = bool(res["base_feat_max_rel_err"] < 1e-4 and res["identical_proposals_frac"] >= 0.98 and
                     res["cls_prob_max_abs_err"] < 1e-4 and res["bbox_pred_max_rel_err"] < 1e-4)
    return res


def run_b200(args):
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner) go to stderr meanwhile
    sys.stdout.flush()
    _saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    from d2t_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the d2t_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # (NCCL_DEBUG is left as the caller set it: fd 1 points at stderr until the JSON line is printed, so NCCL's log
        # lines cannot land on stdout)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    hbm_gbs, _tf, peak_src = peaks()

    net = build_net(101).cuda()
    pairs = PAIRS_PER_GPU
    from d2t_b200.engine import D2TEngine
    engine = D2TEngine(net, pairs, H, W, passes=args.passes)
    im_host, info_host = make_inputs(pairs, seed=1 + rank)          # shard = this rank's own pairs
    im_pin, info_pin = im_host.pin_memory(), info_host.pin_memory()
    im_dev, info_dev = im_pin.cuda(non_blocking=True), info_pin.cuda(non_blocking=True)
    flush = torch.zeros(128 * 1024 * 1024, device="cuda")          # 512 MB

    runner, graph_note = engine, "eager launches (--no-graph)"
    if args.graph:
        # the public serving form: the forward captured once as a CUDA graph, replayed per step.  A capture failure is not
        # a reason to lose the measurement: fall back to eager launches and say so in the output line.
        from d2t_b200.engine import GraphedEngine
        try:
            runner = GraphedEngine(engine, pairs, H, W)
            runner(im_dev, info_dev)
            torch.cuda.synchronize()
            graph_note = "CUDA-graph replay of the eager launch sequence (GraphedEngine)"
        except Exception as exc:   # noqa: BLE001
            print("bench.py: CUDA-graph capture failed (%r); eager launches instead" % (exc,), file=sys.stderr)
            torch.cuda.synchronize()
            runner, graph_note = engine, "eager launches (graph capture failed: %s)" % (repr(exc)[:120],)

    def step(im, info):
        return runner(im, info)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(im_dev, info_dev)
    barrier()

    # ---- device-resident timing: K steps, CUDA events per step, L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    evs = []
    barrier()
    for _ in range(args.steps):
        flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(im_dev, info_dev)
        b.record()
        evs.append((a, b))
    barrier()
    my_launches = ops.LAUNCHES - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    check_finite(step(im_dev, info_dev)[:4], "the eval forward's outputs")
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API: every step copies ITS inputs pinned host -> device, runs the forward and
    # reads ITS results back to pinned host memory.  The copies run on a second stream with double-buffered device
    # inputs, so step i+1's upload overlaps step i's compute (a serving loop); the timed region brackets all K steps
    # including every copy.  No L2 flush here: each step's inputs arrive fresh from the host and the weights +
    # activations (> 3 GB) exceed the 126 MB L2 many times over.
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    dev_in = [(torch.empty_like(im_dev), torch.empty_like(info_dev)) for _ in range(2)]
    up_done = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    outs_pin = None
    d2h = 0

    # the same loop fed with what the reference's loader starts from: uint8 BGR frames as cv2.imread
    # returns them, 720x1200 -> im_scale 600/720 -> 600x1000; the float cast, mean subtraction, bilinear resize and NCHW
    # permute of blob.py / minibatch.py / roibatchLoader.py run on the device (d2t_frames_prep), one launch per step
    from d2t_b200 import synth
    RAW_H, RAW_W = 720, 1200
    raw_hw = ops.frames_resized_shape(RAW_H, RAW_W, 600, 1000, cap=False)
    assert raw_hw[:2] == (H, W), raw_hw
    raw_pin = torch.from_numpy(np.stack([synth.make_frame(RAW_H, RAW_W, 100 + 8 * rank + i) for i in range(2 * pairs)])).pin_memory()
    raw_dev = [torch.empty(raw_pin.shape, dtype=torch.uint8, device="cuda") for _ in range(2)]

    def upload(slot, raw=False):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the forward that last read this slot is done
            if raw:
                raw_dev[slot].copy_(raw_pin, non_blocking=True)
            else:
                dev_in[slot][0].copy_(im_pin, non_blocking=True)
            dev_in[slot][1].copy_(info_pin, non_blocking=True)
            up_done[slot].record(copy_stream)

    def e2e_loop(n, raw=False):
        nonlocal outs_pin, d2h
        upload(0, raw)
        for it in range(n):
            slot = it & 1
            if it + 1 < n:
                upload(slot ^ 1, raw)
            main_stream.wait_event(up_done[slot])
            if raw:                                         # frames in [pair][leg] order = the reference's [pairs][2] batch
                ops.frames_prep(raw_dev[slot], raw_hw[2], out=dev_in[slot][0].view(2 * pairs, 3, H, W))
            o = step(dev_in[slot][0], dev_in[slot][1])[:4]
            consumed[slot].record(main_stream)
            if outs_pin is None:
                outs_pin = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in o]
                d2h = sum(t.numel() * t.element_size() for t in o)
            for dst, src in zip(outs_pin, o):
                dst.copy_(src, non_blocking=True)

    for c in consumed:
        c.record(main_stream)
    e2e_loop(3)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e2e_loop(args.steps)
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    h2d = im_pin.numel() * 4 + info_pin.numel() * 4
    h2d_raw = raw_pin.numel() + info_pin.numel() * 4
    try:                                                    # (a secondary arm: its failure must not cost the headline line;
        e2e_loop(3, raw=True)                               #  every rank still reaches the collectives below)
        torch.cuda.synchronize()
        raw_error = None
    except Exception as exc:   # noqa: BLE001
        raw_error = repr(exc)[:300]
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if raw_error is None:
        e2e_loop(args.steps, raw=True)
    b.record()
    barrier()
    e2e_raw_ms = a.elapsed_time(b) if raw_error is None else float("nan")
    if raw_error is None:
        check_finite([t.cuda() for t in outs_pin], "the outputs of the forward fed with uint8 frames")

    from d2t_b200 import parallel
    t = parallel.max_over_ranks(torch.tensor([total_ms, e2e_ms, e2e_raw_ms], device="cuda", dtype=torch.float64))
    total_ms, e2e_ms, e2e_raw_ms = float(t[0]), float(t[1]), float(t[2])

    # ---- the training step of configs[2] (data-parallel, the path's one collective) on every rank, after the eval numbers
    train = None
    if not args.no_train:
        try:
            train = train_measure(world, rank, steps=5, warmup=3)
        except SystemExit:
            raise
        except Exception as exc:   # noqa: BLE001 -- the eval line must not be lost to the secondary measurement
            train = {"error": repr(exc)[:300]}

    if rank == 0:
        # ---- the dominant kernel: conv_igemm_tf32 (~85 % of the step, profiles/).  All conv launches of one step
        # (trunk + heads, the engine's own layer list) timed together with CUDA events on the launching stream.
        frames = im_dev.permute(1, 0, 2, 3, 4).reshape(2 * pairs, 3, H, W).contiguous()
        from d2t_b200 import conv as dc
        conv_ms = []
        for it in range(3 + 5):
            flush_l2(flush)
            engine.amax.zero()                                    # as every forward does (amax scalars, hand-shake counters)
            engine.stem.run(frames)                               # (stem conv timed too; its pack kernel is not)
            dc.maxpool3x3s2(engine.stem.out, out=engine.pool_out)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for layer in engine.layers:
                layer.run()
            b.record()
            b.synchronize()
            if it >= 3:
                conv_ms.append(a.elapsed_time(b))
        conv_ms = float(np.mean(conv_ms))
        conv_flops = sum(l.flops for l in engine.layers)
        n_conv = len(engine.layers)
        tf_useful = conv_flops / conv_ms / 1e9
        mma_per_flop = 1 if args.passes == 1 else 3
        kind = "f16" if args.passes == 16 else "tf32"
        kind_peak = _tf if args.passes == 16 else _tf / 2.0
        roofline = {"kernel": "conv_igemm (tcgen05 kind::%s implicit GEMM, %d launches/step)" % (kind, n_conv),
                    "bound": "tensor", "achieved": tf_useful, "peak": _tf, "unit": "TFLOP/s", "frac": tf_useful / _tf,
                    "traffic": ncu_traffic().get("conv_igemm", {}).get("dram_bytes_per_launch") if args.passes == 16 else None,
                    "traffic_note": "dram__bytes_read + write per launch, mean over the step's conv launches, ncu (cold cache, "
                                    "serialised: every layer re-reads its input from DRAM); " + str(ncu_traffic().get("_file")),
                    "peak_source": peak_src + " (cuBLAS bf16 burst = the kind::f16 peak at the clocks of this run; kind::tf32 peaks at half of it)",
                    "peak_sustained": peaks.sustained, "frac_vs_sustained_peak": tf_useful / peaks.sustained,
                    "algorithmic_flops_per_launch": conv_flops / n_conv, "avg_launch_ms": conv_ms / n_conv,
                    "conv_ms_per_step": conv_ms,
                    "issued_tensor_tflops": tf_useful * mma_per_flop,
                    "frac_of_kind_peak_issued": tf_useful * mma_per_flop / kind_peak,
                    "note": "achieved = useful fp32-equivalent conv FLOPs; the fp32-accurate split modes (3xFP16 / 3xTF32) "
                            "issue 3 tensor-core MMAs per useful FLOP"}
        ops_bench = op_microbench(flush, hbm_gbs)
        ps = ops_bench["psroi_fwd"]
        roofline_psroi = {"kernel": "psroi_fwd_isat_mc<7, 256 threads, 3 CTAs/SM> (+ psroi_prep) via d2t_psroi_forward", "bound": "hbm",
                          "achieved": ps["gbs"], "peak": hbm_gbs, "unit": "GB/s", "frac": ps["frac_hbm"],
                          "traffic": ncu_traffic().get("psroi_fwd_isat_mc", {}).get("dram_bytes_per_launch"),
                          "traffic_note": "ncu: the features cross HBM once; the 23.5 MB of outputs stay in L2 within the capture",
                          "peak_source": peak_src, "algorithmic_bytes_per_launch": ps["algorithmic_bytes"],
                          "avg_launch_ms": ps["ms"]}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sec, k_run, _ = cpu_path_pairs_per_sec(threads, 2, 1, budget_s=60.0)
            cpu_baseline = {"value": v, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                            "sample": "1 frame-pair (2 frames 600x1000) eval forward x %d on host cores "
                                      "(torch.nn fp32 convs + oracle/ C restatements)" % k_run}
        parity = parity_vs_torch(net, engine, im_dev, info_dev)
        if not parity["ok"]:
            print("bench.py: WARNING parity object outside tolerance: %r" % (parity,), file=sys.stderr)
        ms_per_step = total_ms / args.steps
        config = shared_config()
        line = {"metric": METRIC, "value": parallel.throughput(pairs, ms_per_step, world), "unit": "frame-pairs/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {16: "fp32 (3xFP16 split tensor-core convs: hi/lo fp16 operand pairs, fp32 accumulate, <= 1e-5 of fp64)",
                          3: "fp32 (3xTF32 tensor-core convs, fp32 accumulate)", 1: "tf32"}[args.passes], "data": "synthetic",
                "config": config,
                "detail": {"global_pairs": world * pairs, "parallelism": "dp%d (pairs sharded, no collective)" % world,
                           "convs": engine.conv_backend, "launch": graph_note, "conv_gflop_per_step": engine.conv_flops / 1e9},
                "parity": parity, "train": train,
                "e2e": {"value": world * pairs / (e2e_ms / args.steps / 1e3), "unit": "frame-pairs/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "e2e_raw_frames": {"error": raw_error} if raw_error is not None or not (e2e_raw_ms > 0) else
                                  {"value": world * pairs / (e2e_raw_ms / args.steps / 1e3), "unit": "frame-pairs/s",
                                   "h2d_bytes_per_step": h2d_raw, "d2h_bytes_per_step": d2h,
                                   "source": "uint8 BGR %dx%d frames (pinned host) -> d2t_frames_prep on the device (cast, mean "
                                             "subtraction, OpenCV float32 bilinear resize to %dx%d, NCHW) -> the same forward" % (RAW_H, RAW_W, H, W)},
                "gpu_launches": my_launches, "roofline": roofline, "roofline_psroi": roofline_psroi,
                "cpu_baseline": cpu_baseline, "clocks": clocks,
                "ops": ops_bench}
        sys.stdout.flush()
        os.dup2(_saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def train_measure(world, rank, steps, warmup):
    """one data-parallel TRAINING step per iteration on 2 frame-pairs per GPU -- forward in
    training mode on the tcgen05 engine, target layers + five losses, backward through DgradConv / WgradLayer (tcgen05) and
    the PSRoI / correlation backward kernels, the bucketed gradient all-reduce (NCCL, the path's one collective) overlapped
    with backward on a side stream, SGD, on-device weight re-pack (d2t_b200.train.D2TTrainEngine).  Returns the `train`
    object of the JSON line (device time, max over ranks)."""
    import torch.distributed as dist
    from d2t_b200 import ops, parallel
    from d2t_b200 import synth
    from d2t_b200.train import D2TTrainEngine
    net = build_net(101).cuda()
    pairs = PAIRS_PER_GPU
    im, info = make_inputs(pairs, seed=1 + rank)
    im, info = im.cuda(), info.cuda()
    # a trained trunk's BatchNorm statistics (the reference fine-tunes a pretrained Res-101, resnet.py:304-309): with the
    # identity BN of the random init the activations reach 1e7 after 33 residual blocks and the losses overflow
    synth.calibrate_batchnorm(net, make_inputs(1, seed=1)[0].view(2, 3, H, W).cuda())
    net.train()
    gt = torch.from_numpy(synth.make_gt_boxes(pairs, 30, seed=2 + rank, height=H, width=W)).cuda()
    nb = (gt[..., 4] > 0).sum(-1, keepdim=True)
    engine = D2TTrainEngine(net, pairs, H, W)
    # (fused: torch's single-kernel multi-tensor SGD -- the same update as the reference's torch.optim.SGD, trainval_net.py:340)
    opt = torch.optim.SGD(engine.params, lr=1e-5, momentum=0.9, weight_decay=1e-4, fused=os.environ.get("D2T_SGD_FUSED", "1") != "0")
    n_grad = engine.flat.numel()

    def step():
        out, loss = engine.forward_backward(im, info, gt, nb)
        opt.step()
        engine.refresh_weights()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        loss = step()
    barrier()
    launches0 = ops.LAUNCHES
    comm, exposed = [], []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    barrier()
    ms = a.elapsed_time(b) / steps
    c_ms, e_ms = engine.comm_stats()                               # the last step's buckets
    t = parallel.max_over_ranks(torch.tensor([ms, c_ms, e_ms], device="cuda", dtype=torch.float64))
    ms, c_ms, e_ms = float(t[0]), float(t[1]), float(t[2])
    check_finite([loss.detach()], "the training loss")
    # phase split (rank-local, one extra pass with events between the phases; not part of the timed region)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    with torch.no_grad():
        if engine.g_fwd is not None:
            engine.g_fwd.replay()
        else:
            engine._engine_forward(im, info)
    ev[1].record()
    engine._run_backward()
    ev[2].record()
    torch.cuda.synchronize()
    fwd_ms, bwd_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    conv_flops = sum(l.flops for l in engine.layers) + engine.trk_layer.flops
    return {"metric": "training frame-pairs/sec (Res-101 D&T, 600px; fwd + bwd + gradient all-reduce + SGD)",
            "value": parallel.throughput(pairs, ms, world), "unit": "frame-pairs/s", "ms_per_step": ms, "steps": steps,
            "nranks": world, "pairs_per_gpu": pairs, "loss": float(loss), "loss_finite": bool(torch.isfinite(loss)),
            "collective": "mean all-reduce of %d trainable fp32 gradients (%.0f MB) in %d buckets, NCCL, side stream, issued "
                          "as each bucket's weight gradients are enqueued" % (n_grad, n_grad * 4 / 1e6, len(engine.buckets)),
            "allreduce_ms": c_ms, "allreduce_exposed_ms": e_ms,
            "overlap_frac": (1.0 - e_ms / c_ms) if c_ms > 0 else None,
            "engine_forward_ms": fwd_ms, "engine_backward_ms": bwd_ms,
            "heads_losses_optimizer_ms": max(0.0, ms - fwd_ms - bwd_ms),
            "conv_tflops_useful_fwd": conv_flops / fwd_ms / 1e9,
            "conv_tflops_useful_bwd": (engine.dgrad_flops + engine.wgrad_flops) / bwd_ms / 1e9,
            "convs": "d2t_b200 tcgen05 3xFP16: forward, backward-data (DgradConv) and weight-gradient (WgradLayer); PSRoI, "
                     "correlation, NMS, proposal step: d2t_b200 kernels (forward and backward); losses / target layers / SGD: torch",
            "launch": ("CUDA-graph replays: forward, heads (proposal step + target layers + PSRoI heads + losses + their autograd "
                       "backward, no host round trip), one graph per gradient bucket of the backward pass, weight re-pack; SGD "
                       "eager" if engine.g_heads is not None else
                       ("CUDA-graph replays: forward, backward, weight re-pack; heads eager" if engine.g_fwd is not None
                        else "eager launches")),
            "bn": "calibrated (trained-looking) BatchNorm statistics, frozen", "gpu_launches": ops.LAUNCHES - launches0}


def run_train(args):
    sys.stdout.flush()
    _saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the d2t_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    tr = train_measure(world, rank, args.steps, args.warmup)
    if rank == 0:
        line = dict(tr)
        line.update({"n_gpus": world, "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "dtype": "fp32 (3xFP16 split tensor-core convs)", "data": "synthetic", "mode": "train",
                     "config": {"workload": "Res-101 D&T training step, 600x1000 frame-pairs, 128 RoIs/frame (BASELINE.json "
                                            "configs[2] shape)", "pairs_per_gpu": PAIRS_PER_GPU}})
        sys.stdout.flush()
        os.dup2(_saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement (the `train` object)")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch the forward's kernels one by one instead of replaying them as a CUDA graph "
                         "(d2t_b200.engine.GraphedEngine; measured 5.82 vs 5.89 ms/step)")
    ap.add_argument("--passes", type=int, default=16, choices=[1, 3, 16],
                    help="16 = fp32-accurate fp16-split convolutions (3xFP16, the parity mode, default); "
                         "3 = fp32-accurate 3xTF32; 1 = single-pass TF32")
    ap.add_argument("--ops-only", action="store_true", help="only the per-op microbench (configs[3], [4]); for ncu")
    ap.add_argument("--train", action="store_true",
                    help="secondary number: the data-parallel training step (fwd + bwd + NCCL gradient all-reduce + SGD)")
    args = ap.parse_args()
    if args.ops_only:
        torch.cuda.set_device(0)
        flush = torch.zeros(128 * 1024 * 1024, device="cuda")
        print(json.dumps({"ops": op_microbench(flush, peaks()[0])}), flush=True)
    elif args.impl == "reference":
        run_reference(args)
    elif args.train:
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
