"""D2TEngine -- the eval-mode Detect-to-Track forward (rfcn.py:66-250) on hand-written sm_100a kernels.

It is built FROM a ``model.faster_rcnn.resnet.resnet`` module (same parameters, same state_dict)
for a fixed batch geometry, and computes exactly what ``_RFCN.forward`` computes in eval mode:

  stem 7x7/2 + bn + relu -> maxpool(3,2,ceil) -> layer1..layer4 (frozen-BN bottlenecks, layer4 dilated)
  -> RFCN_net 3x3 dil 6 + relu -> {RFCN_cls_net, RFCN_bbox_net, RPN_Conv -> RPN_cls_score / RPN_bbox_pred}
  -> proposal step (decode + clip + sort + NMS)  -> PSRoI cls / loc + 7x7 vote + softmax
  -> 3 cross-frame correlations -> concat -> corr_bbox_net -> PSRoI tracking + vote.

Every convolution runs on the tcgen05/TMA implicit-GEMM kernel (csrc/conv.cu) with eval-mode
BatchNorm folded into a per-channel scale/shift in the epilogue, the residual add and ReLU fused,
both siamese legs batched as one 2B-image pass (the BN statistics are frozen, so this is the same
function as the reference's python loop over legs).  Activations stay fp32 NHWC between
convs; tensors the reference-layout operators consume (correlation, PSRoI, proposal step) are
emitted as plain fp32 NCHW by the producing conv's epilogue.  ``passes=16`` (default) is the
fp32-accurate fp16-split mode ("3xFP16": hi/lo fp16 operands with per-tensor power-of-two scales,
twice the TF32 tensor rate, correlations included; only the 3-channel stem stays on 3xTF32), ``passes=3`` the
fp32-accurate 3xTF32 mode, ``passes=1`` single-pass TF32.
"""
import os

import torch
import torch.nn.functional as F

from . import conv as dc
from . import ops


def _fold_bn(bn):
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    shift = bn.bias.detach() - bn.running_mean.detach() * scale
    return scale, shift


class D2TEngine(object):
    AMAX_SLOTS = 1024

    def __init__(self, net, pairs, height, width, passes=16, cfg_key="TEST", keep_features=False, private_scratch=True,
                 chain=None):
        """chain: run the trunk / head convolutions as persistent multi-layer launches (dc.ConvChain; 3xFP16 only).
        Bit-identical, but MEASURED SLOWER than per-layer launches with programmatic dependent launch (DESIGN section 6:
        112 against 98 us per layer3 bottleneck), so it is off unless asked for (chain=True or D2T_CONV_CHAIN=1)."""
        dev = next(net.parameters()).device
        self.amax = dc.AmaxArena(self.AMAX_SLOTS, dev)       # every activation tensor's running max |x|, zeroed once per forward
        with self.amax:
            self._build(net, pairs, height, width, passes, cfg_key, keep_features)
        if os.environ.get("D2T_CONV_DONE") == "1":
            # EXPERIMENT: consecutive conv launches hand over through per-layer completion counters instead of
            # griddepcontrol.wait.  The counters live in the amax arena, so the per-forward memset clears them too.
            for chain in (self.layers, self.corr_layers + [self.trk_layer]):
                prev = None
                for layer in chain:
                    layer.set_done(prev, self.amax.take().view(torch.int32))
                    prev = layer
        if private_scratch:    # (default) this engine's conv chain may overlap any other conv work on a different stream;
            # plans WITHOUT a private scratch share one per device and are kept stream-ordered by d2t_conv_plan_run
            from ._lib import lib
            self.scratch = torch.zeros(lib().d2t_conv_scratch_bytes(), dtype=torch.uint8, device=dev)
            for layer in [self.stem, self.trk_layer] + self.layers + self.corr_layers:
                layer.set_scratch(self.scratch)
        if os.environ.get("D2T_CONV_EARLY_B", "1") != "0" and type(self) is D2TEngine:
            for layer in self.layers + [self.trk_layer]:       # weights packed once at build time: nothing upstream writes them
                layer.set_early_weights(True)
        # the engine computes with PACKED copies of the parameters (folded BatchNorm, fp16 operand pairs): remember their
        # version counters so that a later load_state_dict / optimizer step is noticed instead of silently ignored
        self._param_versions = self._versions()
        if chain is None:
            chain = os.environ.get("D2T_CONV_CHAIN", "0") == "1"
        # the launch list of forward(): with chains, maximal runs of plain 3xFP16 layers collapse into one launch each
        self.run_list = dc.build_chains(self.layers) if (chain and passes == 16 and os.environ.get("D2T_CONV_DONE") != "1") \
            else list(self.layers)
        # Two branches after the RPN head (D2T_ENGINE_FORK=0: one stream): the proposal step (decode, sort, NMS mask + sweep:
        # ~0.25 ms of small kernels that leave the device almost empty) on the caller's stream, the remaining convolutions
        # (R-FCN maps, correlations, tracking head) on a side stream; they join at the PSRoI heads.  Captured by
        # GraphedEngine as parallel graph branches.
        self.fork = (os.environ.get("D2T_ENGINE_FORK", "1") != "0" and self.run_list == list(self.layers)
                     and os.environ.get("D2T_CONV_DONE") != "1")
        self.side = torch.cuda.Stream(device=dev) if self.fork else None

    def _build(self, net, pairs, height, width, passes, cfg_key, keep_features):
        from model.utils.config import cfg
        self.net, self.B, self.H, self.W, self.passes = net, pairs, height, width, passes
        # the 3-channel stem: 3xFP16 too (K blocks of two filter rows); D2T_STEM_PASSES=3 keeps the 3xTF32 kernel (A/B runs)
        tf32_passes = passes if passes != 16 else int(os.environ.get("D2T_STEM_PASSES", "16"))
        self.cfg_key = cfg_key
        self.keep_features = keep_features   # also emit conv3/4/5 as plain NCHW (tests / inspection)
        self.post_nms = cfg[cfg_key].RPN_POST_NMS_TOP_N
        self.pre_nms = cfg[cfg_key].RPN_PRE_NMS_TOP_N
        self.nms_thresh = cfg[cfg_key].RPN_NMS_THRESH
        self.n_classes, self.n_reg = net.n_classes, net.n_reg_classes
        dev = next(net.parameters()).device
        self.device = dev
        N = 2 * pairs
        self.N = N
        self.layers = []          # run order
        self.conv_flops = 0.0
        base = net.RFCN_base

        # ---- stem + pool
        s, b = _fold_bn(base[1])
        self.stem = dc.StemConv(N, height, width, base[0].weight, s, b, relu=True, passes=tf32_passes, device=dev)
        self.conv_flops += self.stem.flops
        so = self.stem.out
        ph = -(-(so.H - 3) // 2) + 1
        pw = -(-(so.W - 3) // 2) + 1
        ph -= 1 if (ph - 1) * 2 >= so.H else 0
        pw -= 1 if (pw - 1) * 2 >= so.W else 0
        self.pool_out = dc.ActTensor(N, ph, pw, 64, device=dev, amax=so.amax)
        x = self.pool_out

        # ---- residual stages
        self.feat_nchw, self.feat_nhwc = {}, {}
        for idx in (4, 5, 6, 7):
            stage = base[idx]
            for bi, blk in enumerate(stage):
                last = bi == len(stage) - 1
                x = self._bottleneck(x, blk, feature=(last and idx in (5, 6, 7)), tag=idx)
        self.conv5 = x
        # ---- head conv (3x3 dilation 6, bias) + relu
        rn = base.RFCN_net
        self.base_feat = self._conv(x, rn.weight, None, rn.bias, 1, rn.padding[0], rn.dilation[0], relu=True).out
        bf = self.base_feat
        # ---- RPN head (first: the proposal step that consumes it is a chain of small latency-bound kernels, which forward()
        # overlaps with everything below)
        rpn = net.RFCN_rpn
        rc = self._conv(bf, rpn.RPN_Conv.weight, None, rpn.RPN_Conv.bias, 1, 1, 1, relu=True).out
        self.rpn_score = self._conv(rc, rpn.RPN_cls_score.weight, None, rpn.RPN_cls_score.bias, want_nhwc=False,
                                    want_nchw=True).out_nchw
        self.rpn_delta = self._conv(rc, rpn.RPN_bbox_pred.weight, None, rpn.RPN_bbox_pred.bias, want_nhwc=False,
                                    want_nchw=True).out_nchw
        self.n_main_layers = len(self.layers)     # layers[:n_main_layers] feed the proposal step; the rest only the PSRoI heads
        # ---- R-FCN maps: NCHW for PSRoI; the loc map also feeds the tracking concat (NHWC slices)
        cn, bn_ = net.RFCN_cls_net, net.RFCN_bbox_net
        self.cls_map = self._conv(bf, cn.weight, None, cn.bias, want_nhwc=False, want_nchw=True).out_nchw
        n_loc = 4 * self.n_reg * 49
        c3c, c45c = 81, 289
        self.trk_cin = 2 * n_loc + c3c + 2 * c45c
        self.trk_in = dc.ActTensor(pairs, bf.H, bf.W, self.trk_cin, cstride=(self.trk_cin + 31) // 32 * 32, device=dev)
        self.bbox_map = torch.empty(N, n_loc, bf.H, bf.W, device=dev)
        for leg in (0, 1):   # one plan per leg: NHWC into its channel slice of the concat buffer + NCHW for PSRoI
            self._conv(bf.batch_slice(leg * pairs, (leg + 1) * pairs), bn_.weight, None, bn_.bias, out=self.trk_in,
                       out_coffset=leg * n_loc, out_nchw=self.bbox_map[leg * pairs:(leg + 1) * pairs])
        self.n_trunk_layers = len(self.layers)
        # ---- tracking head conv (runs after the correlations)
        tn = net.corr_bbox_net
        self.trk_layer = self._make_layer(self.trk_in, tn.weight, None, tn.bias, want_nhwc=False, want_nchw=True)
        self.conv_flops += self.trk_layer.flops
        self.anchors = rpn.RPN_proposal._anchors.to(dev)
        self.zero_losses = tuple(torch.zeros(2, 1, device=dev) for _ in range(4))
        self.zero_track_loss = torch.zeros(1, device=dev)
        self.feat_stride = rpn.feat_stride
        # ---- correlations: straight from the NHWC split features of the two legs into the concat buffer
        self.corr_layers = []
        for corr, tag, coff in ((net.conv3_corr_layer, 5, 2 * n_loc), (net.conv4_corr_layer, 6, 2 * n_loc + c3c),
                                (net.conv5_corr_layer, 7, 2 * n_loc + c3c + c45c)):
            f = self.feat_nhwc[tag]
            assert corr.kernel_size == 1 and corr.stride1 == corr.stride2
            self.corr_layers.append(dc.CorrLayer(f.batch_slice(0, pairs), f.batch_slice(pairs, N), corr.pad_size,
                                                 corr.max_displacement, corr.stride1, passes=passes, out=self.trk_in,
                                                 out_coffset=coff))
        kind = {16: "kind::f16 x3 passes (fp16 hi/lo split, per-tensor 2^k scales)", 3: "kind::tf32 x3 passes",
                1: "kind::tf32 x1 pass"}[passes]
        self.conv_backend = "d2t_b200 tcgen05 implicit GEMM, %s, TMA-fed, fused BN/ReLU/residual" % kind

    def _versions(self):
        return tuple(t._version for t in list(self.net.parameters()) + list(self.net.buffers()))

    def check_fresh(self):
        """raise if a parameter or BatchNorm buffer of the module was modified in place after the engine packed it
        (load_state_dict, an optimizer step, .data.copy_): rebuild the engine -- or use D2TTrainEngine, which re-packs"""
        if self._param_versions is not None and self._versions() != self._param_versions:
            raise RuntimeError("D2TEngine: the module's parameters changed after the engine was built (it computes with packed "
                               "copies); build a new engine after load_state_dict / parameter updates")

    # ------------------------------------------------------------------ construction helpers
    def _make_layer(self, x, weight, scale, shift, stride=1, pad=0, dil=1, relu=False, residual=None, **kw):
        """the one place a convolution plan is created (the training engine overrides it)"""
        return dc.ConvLayer(x, weight, scale, shift, stride, pad, dil, relu, residual, passes=self.passes, **kw)

    def _conv(self, x, weight, scale, shift, stride=1, pad=0, dil=1, relu=False, residual=None, **kw):
        layer = self._make_layer(x, weight, scale, shift, stride, pad, dil, relu, residual, **kw)
        self.layers.append(layer)
        self.conv_flops += layer.flops
        return layer

    def _bottleneck(self, x, blk, feature, tag):
        want_nchw = feature and self.keep_features
        s1, b1 = _fold_bn(blk.bn1)
        s2, b2 = _fold_bn(blk.bn2)
        s3, b3 = _fold_bn(blk.bn3)
        c1, c2, c3 = blk.conv1, blk.conv2, blk.conv3
        if blk.downsample is not None:
            sd, bd = _fold_bn(blk.downsample[1])
            dconv = blk.downsample[0]
            res = self._conv(x, dconv.weight, sd, bd, dconv.stride[0]).out
        else:
            res = x
        y = self._conv(x, c1.weight, s1, b1, c1.stride[0], relu=True).out
        y = self._conv(y, c2.weight, s2, b2, 1, c2.padding[0], c2.dilation[0], relu=True).out
        last = self._conv(y, c3.weight, s3, b3, relu=True, residual=res, want_nchw=want_nchw)
        if feature:
            self.feat_nhwc[tag] = last.out
            if want_nchw:
                self.feat_nchw[tag] = last.out_nchw
        return last.out

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, im_data, im_info):
        """im_data [B, 2, 3, H, W], im_info [B, 2, 3] (CUDA fp32) -> the reference's 10-tuple (eval)."""
        self.check_fresh()
        info = self._begin(im_data, im_info)
        if not self.fork:
            for item in self.run_list:
                item.run()
            return self._tail(im_data, info)
        for layer in self.layers[:self.n_main_layers]:
            layer.run()
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        h = self.side.cuda_stream
        for layer in self.layers[self.n_main_layers:] + self.corr_layers + [self.trk_layer]:
            layer.run(h)
        return self._tail(im_data, info, joined=self.side)

    def _begin(self, im_data, im_info):
        """input re-layout, stem conv and max-pool; returns the per-frame im_info [2B, 3]"""
        N = self.N
        assert tuple(im_data.shape) == (self.B, 2, 3, self.H, self.W), "engine was built for a fixed geometry"
        info = im_info.permute(1, 0, 2).reshape(N, 3).contiguous().float()
        self.amax.zero()
        if self.stem.passes == 16 and im_data.is_contiguous():
            self.stem.run(im_data, pairs=self.B)               # (the leg-major order is the packer's read pattern)
        else:
            self.stem.run(im_data.permute(1, 0, 2, 3, 4).reshape(N, 3, self.H, self.W).contiguous())
        dc.maxpool3x3s2(self.stem.out, out=self.pool_out)
        return info

    def _tail(self, im_data, info, joined=None):
        """everything after the trunk / head convs: proposal step, PSRoI heads, correlations, tracking head.
        joined: the side stream on which forward() has already enqueued the R-FCN map convolutions, the correlations and
        the tracking head; it is waited for before the first PSRoI head."""
        B, L, N = self.B, 2, self.N
        # ---- proposals for all 2B images
        A = self.anchors.size(0)
        sc = self.rpn_score
        prob = F.softmax(sc.view(N, 2, A * sc.size(2), sc.size(3)), dim=1).view_as(sc)   # rpn.py:66-68
        rois_all = ops.proposals(self.anchors, self.rpn_delta, prob, info, self.feat_stride, self.pre_nms,
                                 self.post_nms, self.nms_thresh)                        # [2B, R, 5]
        R = rois_all.size(1)
        flat = rois_all.view(-1, 5)
        if joined is not None:
            torch.cuda.current_stream().wait_stream(joined)
        # ---- detection heads: PSRoI pooling + 7x7 vote (+ softmax) fused (rfcn.py:133-140)
        cls_prob = ops.psroi_vote(self.cls_map, flat, 7, 7, 1.0 / 16.0, 7, self.n_classes, softmax=True).view(L, B, R, -1)
        bbox_pred = ops.psroi_vote(self.bbox_map, flat, 7, 7, 1.0 / 16.0, 7, 4 * self.n_reg).view(L, B, R, -1)
        rois = rois_all.view(L, B, R, 5).clone()
        rois[1, :, :, 0] -= B                                                           # per-leg image index
        # ---- tracking branch
        if joined is None:
            for layer in self.corr_layers:
                layer.run()
            self.trk_layer.run()
        tracking_pred = ops.psroi_vote(self.trk_layer.out_nchw, rois[0].reshape(-1, 5), 7, 7, 1.0 / 16.0, 7,
                                       4 * self.n_reg).view(B * R, -1)                  # rfcn.py:192-196
        # (the four eval-mode losses and the tracking loss: constant zeros, allocated once -- five fill / copy launches fewer
        # at the serial end of the step)
        return (rois, cls_prob, bbox_pred, tracking_pred) + self.zero_losses + ([], self.zero_track_loss)

    __call__ = forward

    @torch.no_grad()
    def detect(self, im_data, im_info, thresh=0.0, nms_thresh=None):
        """forward + the reference's post-processing (test_net.py:232-301) for every frame and class in one batched
        pass: returns (forward outputs, d2t_b200.detect.Detections)."""
        return _detect(self, self.n_reg, im_data, im_info, thresh, nms_thresh)


@torch.no_grad()
def _detect(engine, n_reg, im_data, im_info, thresh, nms_thresh):
    from model.utils.config import cfg
    from . import detect as det
    out = engine(im_data, im_info)
    nms_t = cfg.TEST.NMS if nms_thresh is None else nms_thresh
    return out, det.per_class_detections(out[0], out[1], out[2], im_info, thresh, nms_t, class_agnostic=(n_reg == 1),
                                         stds=cfg.TRAIN.BBOX_NORMALIZE_STDS, means=cfg.TRAIN.BBOX_NORMALIZE_MEANS)


class D2TEngineStreams(object):
    """The same forward as ``D2TEngine(net, pairs, ...)`` run as ``chains`` independent sub-engines (contiguous groups
    of frame-pairs, SURVEY 8e: pairs are independent units) on their own CUDA streams, ENQUEUED LAYER BY LAYER IN TURN.
    Every conv launch of one chain depends on the launch before it, and that dependency costs a fixed ~7.5 us of idle
    SMs per launch (finding 4); with two chains in flight the gap of one hides behind the other's
    kernel.  Each sub-engine owns a private stream-K scratch.  The outputs are concatenated in the reference's layout
    with the image index in column 0 of ``rois`` counted over the whole batch, exactly as ``D2TEngine`` returns them."""

    def __init__(self, net, pairs, height, width, chains=2, **kw):
        assert pairs % chains == 0, "pairs must split evenly over the chains"
        self.B, self.chains, self.per = pairs, chains, pairs // chains
        self.engines = [D2TEngine(net, self.per, height, width, private_scratch=True, **kw) for _ in range(chains)]
        self.streams = [torch.cuda.Stream() for _ in range(chains)]
        e0 = self.engines[0]
        self.conv_flops = sum(e.conv_flops for e in self.engines)
        self.conv_backend = e0.conv_backend + ", %d interleaved chains" % chains
        self.layers = [l for e in self.engines for l in e.layers]
        self.n_classes, self.n_reg = e0.n_classes, e0.n_reg

    @torch.no_grad()
    def forward(self, im_data, im_info):
        cur = torch.cuda.current_stream()
        per = self.per
        infos, outs = [], []
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                infos.append(e._begin(im_data[i * per:(i + 1) * per], im_info[i * per:(i + 1) * per]))
        handles = [s.cuda_stream for s in self.streams]
        for layers in zip(*[e.layers for e in self.engines]):
            for layer, h in zip(layers, handles):
                layer.run(h)
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            with torch.cuda.stream(s):
                o = e._tail(im_data[i * per:(i + 1) * per], infos[i])
                if i:
                    o[0][..., 0] += i * per          # image index inside the whole batch
                outs.append(o)
        for s in self.streams:
            cur.wait_stream(s)
        rois = torch.cat([o[0] for o in outs], 1)
        cls_prob = torch.cat([o[1] for o in outs], 1)
        bbox_pred = torch.cat([o[2] for o in outs], 1)
        tracking_pred = torch.cat([o[3] for o in outs], 0)
        o0 = outs[0]
        return (rois, cls_prob, bbox_pred, tracking_pred) + tuple(o0[4:])

    __call__ = forward

    def detect(self, im_data, im_info, thresh=0.0, nms_thresh=None):
        return _detect(self, self.n_reg, im_data, im_info, thresh, nms_thresh)


class GraphedEngine(object):
    """One forward of a ``D2TEngine`` / ``D2TEngineStreams`` captured as a CUDA graph (streams forked inside the capture
    become parallel branches) and replayed per call: no per-launch host work, and the branches' kernels are released by
    the device, not by the order in which Python reached them.  Inputs are copied into static buffers; the returned
    tensors are the graph's static outputs and are overwritten by the next call.  Safe to replay because the conv
    kernel's stream-K hand-shake is self-cleaning (csrc/conv.cu: the finisher clears the flags it consumed)."""

    def __init__(self, engine, pairs, height, width, device="cuda", warmup=3):
        self.engine = engine
        self.im_data = torch.zeros(pairs, 2, 3, height, width, device=device)
        self.im_info = torch.tensor([float(height), float(width), 1.0], device=device).view(1, 1, 3).repeat(pairs, 2, 1)
        self.graph = None
        self.warmup = warmup
        self.conv_flops = engine.conv_flops
        self.conv_backend = engine.conv_backend + ", CUDA-graph replay"
        self.layers = engine.layers

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.engine(self.im_data, self.im_info)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        count = ops.LAUNCHES
        with torch.cuda.graph(self.graph):
            self.out = self.engine(self.im_data, self.im_info)
        self.launches_per_replay = ops.LAUNCHES - count

    @torch.no_grad()
    def forward(self, im_data, im_info):
        self.im_data.copy_(im_data, non_blocking=True)
        self.im_info.copy_(im_info, non_blocking=True)
        for eng in getattr(self.engine, "engines", [self.engine]):
            eng.check_fresh()
        if self.graph is None:
            self._capture()
        self.graph.replay()
        ops._count(self.launches_per_replay)
        return self.out

    __call__ = forward

    def detect(self, im_data, im_info, thresh=0.0, nms_thresh=None):
        """graph replay + the batched post-processing of D2TEngine.detect (eager: its output sizes depend on the data)"""
        return _detect(self, self.engine.n_reg, im_data, im_info, thresh, nms_thresh)
