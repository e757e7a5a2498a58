"""D2TTrainEngine -- one data-parallel TRAINING step of the Detect-to-Track graph on the hand-written sm_100a kernels.

The reference's step (trainval_net.py:355-373): forward in training mode (rfcn.py:66-250 with the RPN / RCNN / tracking
target layers and five losses), ``loss.backward()`` through every trainable convolution above layer1
(resnet.py:279-295: stem + layer1 and every BatchNorm are frozen), gradient exchange (nn.DataParallel's reduce,
trainval_net.py:310-311), SGD.  Here:

  forward   the engine of d2t_b200.engine (tcgen05 implicit-GEMM convolutions, tensor-core correlations); every
            activation stays resident for the backward pass; weights are re-packed on the device after each update
  heads     proposal step + NMS, target layers, PSRoI pooling (forward / backward kernels), votes and the five
            losses through torch autograd on LEAF copies of the five convolution outputs they consume
            (model/rpn/proposal_target_layer_cascade.py: train_heads) -- exactly the reference's arithmetic
  backward  explicit, layer by layer, in reverse: for y = relu(scale * conv(x, w) + shift [+ residual])
              g      = dL/d(pre-activation), masked by y > 0 in the epilogue of the kernel that produced it
              dw     = scale * x (*) g            WgradLayer  (tcgen05 GEMM with K over the pixels)
              dx    += conv(g, flip(w)^T * scale)  DgradConv   (the forward kernel on the transposed filter; the
                                                              skip connection's gradient enters as its residual)
            correlation backward = the exact-adjoint kernels of csrc/correlation.cu
  exchange  the gradients live in ONE flat buffer in backward order; each bucket is all-reduced (mean) on a side
            stream as soon as the weight gradients that fill it have been enqueued, overlapping the rest of backward
  update    torch.optim.SGD on the same parameters (param.grad are views of the flat buffer)
"""
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import conv as dc
from . import ops
from .engine import D2TEngine, _fold_bn


def _is_param(t):
    return isinstance(t, nn.Parameter)


class D2TTrainEngine(D2TEngine):
    AMAX_SLOTS = 4096

    def __init__(self, net, pairs, height, width, bucket_bytes=32 << 20, use_graphs=True, graph_heads=True):
        self.use_graphs, self.g_fwd, self.g_heads, self.graph_heads = use_graphs, None, None, graph_heads
        self._widx, self._blocks, self._meta = {}, [], {}
        dev = next(net.parameters()).device
        self.w_amax = torch.zeros(512, device=dev)          # max |w| per distinct conv weight (slot = _widx[id(w)])
        self.wt_amax = torch.zeros(512, device=dev)         # x max |folded BN scale|: bound for the backward-data operand
        self._smax = torch.ones(512, device=dev)
        self._weights = []
        D2TEngine.__init__(self, net, pairs, height, width, passes=16, cfg_key="TRAIN", keep_features=True)
        self._param_versions = None        # (this engine re-packs its weights after every optimizer step: refresh_weights)
        self.bucket_bytes = bucket_bytes
        with self.amax:
            self._build_backward()
        self._repack_many = dc.RepackMany([self.trk_layer] + self.layers + self.dgrads)
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.ready = torch.cuda.Event()
        self.allreduce_ms = None

    # ------------------------------------------------------------------ forward construction hooks
    def _slot(self, weight, scale):
        key = id(weight)
        if key not in self._widx:
            i = len(self._weights)
            self._widx[key] = i
            self._weights.append(weight)
            self.w_amax[i] = weight.detach().abs().max()
            self._smax[i] = scale.detach().abs().max() if scale is not None else 1.0
            self.wt_amax[i] = self.w_amax[i] * self._smax[i]
        return self._widx[key]

    def _make_layer(self, x, weight, scale, shift, stride=1, pad=0, dil=1, relu=False, residual=None, **kw):
        i = self._slot(weight, scale)
        layer = dc.ConvLayer(x, weight, scale, shift, stride, pad, dil, relu, residual, passes=16,
                             amax_w=self.wt_amax[i:i + 1], **kw)          # (bound of |w * scale|: the forward pack folds it too)
        layer.meta = dict(weight=weight, scale=scale, bias=shift if _is_param(shift) else None, stride=stride, pad=pad,
                          dil=dil, relu=relu, slot=i)
        self._meta.setdefault(id(weight), []).append(layer)
        return layer

    def _bottleneck(self, x, blk, feature, tag):
        n0 = len(self.layers)
        out = D2TEngine._bottleneck(self, x, blk, feature, tag)
        new = self.layers[n0:]
        self._blocks.append(dict(x=x, ds=new[0] if len(new) == 4 else None, c1=new[-3], c2=new[-2], c3=new[-1], tag=tag,
                                 trainable=blk.conv1.weight.requires_grad))
        return out

    # ------------------------------------------------------------------ backward construction
    def _G(self, like, C=None):
        """gradient buffer shaped like activation `like` (channel stride padded for the 64-channel K blocks)"""
        C = like.C if C is None else C
        return dc.ActTensor(like.N, like.H, like.W, C, cstride=(C + 31) // 32 * 32, device=self.device)

    def _build_backward(self):
        net, B, N, dev = self.net, self.B, self.N, self.device
        params = [p for p in net.parameters() if p.requires_grad]
        self.params = params
        self.flat = torch.zeros(sum(p.numel() for p in params), device=dev)
        self._flat_used = 0
        self._grad_of = {}
        self.bwd = []                # (callable, flat offset reached once it has been enqueued)
        self.dgrads = []
        bf = self.base_feat
        H, W = bf.H, bf.W
        n_loc = 4 * self.n_reg * 49
        layer_of = lambda w: self._meta[id(w)][0]
        # wgrad plane scratch: sized for the largest (input planes, gradient planes) over all trainable convolutions
        need_x, need_g = 0, 0
        for l in self.layers + [self.trk_layer]:
            m = l.meta
            if not m["weight"].requires_grad:
                continue
            O, I, R, S = m["weight"].shape
            oh, ow = l.info["OH"], l.info["OW"]
            nimg = N if l is not self.trk_layer else B
            need_x = max(need_x, nimg * I * (oh if m["stride"] > 1 else l.x.H) * dc._pad32(ow if m["stride"] > 1 else l.x.W))
            need_g = max(need_g, S * nimg * O * oh * dc._pad32(ow))
        # D2T_WGRAD_PREPACK=0: every weight-gradient layer packs its input planes inside the backward pass (one shared
        # buffer); default: own planes per layer (+ ~3 GB), packed on a side stream while the heads run (backward 14.5 ->
        # 13.5 ms, heads 4.3 -> 4.8 ms: step 24.06 -> 23.6 ms; capping the packer's blocks per SM so that the heads' small
        # kernels find room only made the heads wait for the packs: measured worse at every setting)
        self.prepack = os.environ.get("D2T_WGRAD_PREPACK", "1") != "0"
        self.pack_stream = torch.cuda.Stream(device=dev)
        self.g_xpack = None
        self.wscratch = dc.WgradScratch(need_x if not self.prepack else 1, need_g, dev)

        # ---- gradient leaves coming out of the autograd part (NCHW -> NHWC each step)
        cn, bn_, rpn, tn = net.RFCN_cls_net, net.RFCN_bbox_net, net.RFCN_rpn, net.corr_bbox_net
        rc = layer_of(rpn.RPN_Conv.weight).out
        self.g_cls = self._G(bf, self.n_classes * 49)
        self.g_bbox = self._G(bf, n_loc)
        self.g_score = self._G(bf, rpn.nc_score_out)
        self.g_delta = self._G(bf, rpn.nc_bbox_out)
        self.g_trk = dc.ActTensor(B, H, W, n_loc, cstride=dc._pad32(n_loc), device=dev)
        self.g_trk_in = dc.ActTensor(B, H, W, self.trk_in.cstride, cstride=self.trk_in.cstride, device=dev)

        # ---- correlation backward: (gradient slice of the concat buffer, other frame's features) -> feature gradients
        self.corr_bwd, self._extra = [], {}
        coff = 2 * n_loc
        specs = []
        for corr, tag in ((net.conv3_corr_layer, 5), (net.conv4_corr_layer, 6), (net.conv5_corr_layer, 7)):
            f = self.feat_nhwc[tag]
            st, r = corr.stride1, corr.max_displacement // corr.stride2
            assert corr.kernel_size == 1 and corr.stride1 == corr.stride2 and corr.pad_size == corr.max_displacement
            specs.append((corr, tag, f, st, r, coff))
            coff += (2 * r + 1) ** 2
        nb_, no_ = 0, 0
        for corr, tag, f, st, r, co in specs:
            a, b_ = dc.CorrBwdScratch.need(B, f.C, -(-f.H // st), -(-f.W // st), r)
            nb_, no_ = max(nb_, a), max(no_, b_)
        self.cscratch = dc.CorrBwdScratch(nb_, no_, dev)
        for corr, tag, f, st, r, co in specs:
            # the gradient lives on the correlation lattice: for conv3 (stride 2) that is the even positions, a
            # [N, 38, 63, C] tensor that joins the stride-2 backward-data of layer3's first block before the scatter
            ex = dc.ActTensor(N, -(-f.H // st), -(-f.W // st), f.C, cstride=f.C, device=dev)
            self._extra[tag] = ex
            for which in (1, 2):
                other = f.batch_slice(B, N) if which == 1 else f.batch_slice(0, B)
                out = ex.batch_slice(0, B) if which == 1 else ex.batch_slice(B, N)
                layer = dc.CorrBwdLayer(self.g_trk_in, co, other, out, corr.max_displacement, st, which, self.cscratch)
                layer.set_scratch(self.scratch)
                self.corr_bwd.append(layer)
        # ---- tracking head conv
        self._wgrad(self.trk_layer, self.trk_in, self.g_trk)
        self._dgrad(self.trk_layer, self.g_trk, self.g_trk_in, out_channels=self.trk_in.cstride)
        self.bwd.append((self._tracking_split, None, 'tracking split + correlation backward'))
        # ---- heads on base_feat
        self._wgrad(layer_of(cn.weight), bf, self.g_cls)
        self._wgrad(layer_of(bn_.weight), bf, self.g_bbox)
        self._wgrad(layer_of(rpn.RPN_cls_score.weight), rc, self.g_score)
        self._wgrad(layer_of(rpn.RPN_bbox_pred.weight), rc, self.g_delta)
        g_rc = self.g_rc = self._G(rc)
        self._dgrad(layer_of(rpn.RPN_cls_score.weight), self.g_score, g_rc, mask=rc)
        self._dgrad(layer_of(rpn.RPN_bbox_pred.weight), self.g_delta, g_rc, residual=g_rc, mask=rc)
        self._wgrad(layer_of(rpn.RPN_Conv.weight), bf, g_rc)
        g_base = self.g_base = self._G(bf)
        self._dgrad(layer_of(cn.weight), self.g_cls, g_base, mask=bf)
        self._dgrad(layer_of(bn_.weight), self.g_bbox, g_base, residual=g_base, mask=bf)
        self._dgrad(layer_of(rpn.RPN_Conv.weight), g_rc, g_base, residual=g_base, mask=bf)
        # ---- head conv on conv5
        head = layer_of(net.RFCN_base.RFCN_net.weight)
        self._wgrad(head, self.conv5, g_base)
        g = self.g_conv5 = self._G(self.conv5)
        self._dgrad(head, g_base, g, residual=self.extra[7], mask=self.conv5)
        # ---- residual stages, last block first
        for bi in range(len(self._blocks) - 1, -1, -1):
            rec = self._blocks[bi]
            if not rec["trainable"]:
                break
            first_of_stage = bi == 0 or self._blocks[bi - 1]["tag"] != rec["tag"]
            need_dx = self._blocks[bi - 1]["trainable"] if bi > 0 else False
            extra = self.extra.get(rec["tag"] - 1) if first_of_stage else None
            g = self._bottleneck_bwd(rec, g, need_dx, extra)
        assert self._flat_used == self.flat.numel(), (self._flat_used, self.flat.numel())
        # ---- buckets: contiguous ranges of the flat gradient buffer, closed in backward order
        self.buckets, start = [], 0
        for k, (_, off, _l) in enumerate(self.bwd):
            if off is not None and ((off - start) * 4 >= self.bucket_bytes or off == self.flat.numel()):
                self.buckets.append((k, start, off))
                start = off
        for p in params:
            p.grad = self._grad_of[id(p)]

    @property
    def extra(self):
        """gradients arriving at the conv3 / conv4 / conv5 features from the three correlations (tag -> ActTensor on the
        correlation lattice)"""
        return self._extra

    def _grad_buf(self, p):
        if id(p) not in self._grad_of:
            n = p.numel()
            self._grad_of[id(p)] = self.flat[self._flat_used:self._flat_used + n].view_as(p)
            self._flat_used += n
        return self._grad_of[id(p)]

    def _wgrad(self, layer, x, g):
        m = layer.meta
        w = m["weight"]
        if not w.requires_grad:
            return
        gw = self._grad_buf(w)
        scale = m["scale"]
        wl = dc.WgradLayer(x, g, gw, scale, m["stride"], m["pad"], m["dil"], self.wscratch, own_xt=self.prepack)
        wl.set_scratch(self.scratch)
        self.wgrads = getattr(self, "wgrads", []) + [wl]
        self.wgrad_flops = getattr(self, "wgrad_flops", 0.0) + wl.flops
        steps = [wl.run]
        if m["bias"] is not None:
            gb = self._grad_buf(m["bias"])
            C_ = gb.numel()
            flat2d = g.x.view(-1, g.cstride)
            steps.append(lambda: torch.sum(flat2d[:, :C_], 0, out=gb))
        off = self._flat_used
        self.bwd.append((lambda: [s() for s in steps], off, 'wgrad %dx%d %d->%d @%dx%d' % (gw.shape[2], gw.shape[3], gw.shape[1], gw.shape[0], g.H, g.W)))

    def _dgrad(self, layer, g, out, residual=None, mask=None, out_channels=None):
        m = layer.meta
        i = m["slot"]
        d = dc.DgradConv(g, m["weight"], m["scale"], m["pad"], m["dil"], self.wt_amax[i:i + 1], out=out, residual=residual,
                         mask=mask, out_channels=out_channels)
        d.set_scratch(self.scratch)
        self.dgrads.append(d)
        self.dgrad_flops = getattr(self, "dgrad_flops", 0.0) + d.flops
        self.bwd.append((d.run, None, 'dgrad %dx%d %d->%d @%dx%d' % (m['weight'].shape[2], m['weight'].shape[3], m['weight'].shape[0], m['weight'].shape[1], g.H, g.W)))
        return d

    def _bottleneck_bwd(self, rec, g_out, need_dx, extra):
        """g_out: gradient w.r.t. the block output's pre-activation (already masked by the output's ReLU).  Returns the
        same for the block's input (None if nobody needs it)."""
        x, c1, c2, c3, ds = rec["x"], rec["c1"], rec["c2"], rec["c3"], rec["ds"]
        self._wgrad(c3, c2.out, g_out)
        g2 = self._G(c2.out)
        self._dgrad(c3, g_out, g2, mask=c2.out)
        self._wgrad(c2, c1.out, g2)
        g1 = self._G(c1.out)
        self._dgrad(c2, g2, g1, mask=c1.out)
        self._wgrad(c1, x, g1)
        if ds is not None:
            self._wgrad(ds, x, g_out)
        if not need_dx:
            return None
        stride = c1.meta["stride"]
        gx = self._G(x)
        if stride == 1:
            if ds is None:
                self._dgrad(c1, g1, gx, residual=g_out, mask=x)          # + the skip connection's gradient
            else:
                self._dgrad(c1, g1, gx, residual=extra, mask=x)
                self._dgrad(ds, g_out, gx, residual=gx, mask=x)
        else:
            assert stride == 2 and ds is not None
            low = dc.ActTensor(g1.N, g1.H, g1.W, x.C, cstride=gx.cstride, device=self.device)
            self._dgrad(c1, g1, low, residual=extra)                     # (the stride-2 correlation's gradient: same lattice)
            self._dgrad(ds, g_out, low, residual=low)
            assert gx.cstride == x.cstride == x.C, "the scatter kernel wants one channel stride"
            self.bwd.append((lambda: dc.upsample2_add_mask(low, gx, mask=x), None, 'upsample2_add_mask'))
        return gx

    # ------------------------------------------------------------------ per-step pieces
    def _tracking_split(self):
        """g_trk_in [B, H, W, 2*n_loc + 81 + 289 + 289] -> the two legs of the loc map's gradient and, through the
        tensor-core correlation backward (csrc/conv.cu, CORRB), the gradients of the conv3 / conv4 / conv5 features"""
        B = self.B
        n_loc = 4 * self.n_reg * 49
        gi = self.g_trk_in
        self.g_bbox.x[:B, :, :, :n_loc] += gi.x[..., :n_loc]
        self.g_bbox.x[B:, :, :, :n_loc] += gi.x[..., n_loc:2 * n_loc]
        torch.amax(self.g_bbox.x.abs().view(-1), 0, keepdim=True, out=self.g_bbox.amax)
        for layer in self.corr_bwd:
            layer.run()

    def refresh_weights(self):
        """after an optimizer step: new max |w| per weight (one fused norm), then the packed fp16 operand pairs of every
        forward and backward-data plan -- all on the device, no host synchronisation"""
        if self.g_fwd is not None:
            self.g_refresh.replay()
            ops._count(self._refresh_launches)
        else:
            with torch.no_grad():
                self._refresh_weights()

    def _refresh_weights(self):
        norms = torch._foreach_norm([w.detach() for w in self._weights], float("inf"))
        n = len(norms)
        torch.stack(norms, out=self.w_amax[:n])
        torch.mul(self.w_amax[:n], self._smax[:n], out=self.wt_amax[:n])
        self._repack_many.run()                   # one launch for all ~630 forward / backward-data operands

    # ------------------------------------------------------------------ the step
    def _engine_forward(self, im_data, im_info):
        info = self._begin(im_data, im_info)
        for layer in self.layers:
            layer.run()
        for layer in self.corr_layers:
            layer.run()
        self.trk_layer.run()
        return info

    def _prepack_inputs(self):
        """the input planes of every weight-gradient GEMM (needs the forward activations only)"""
        for wl in self.wgrads:
            wl.pack_input()

    def _load_leaf_grads(self, grads):
        for buf, gr in zip((self.g_cls, self.g_bbox, self.g_score, self.g_delta, self.g_trk), grads):
            buf.load_nchw(gr)

    def _segments(self):
        """backward steps grouped so that a gradient bucket closes at the end of each group"""
        cuts = [k + 1 for k, _, _ in self.buckets]
        if not cuts or cuts[-1] != len(self.bwd):
            cuts.append(len(self.bwd))
        return list(zip([0] + cuts[:-1], cuts))

    def _capture(self, im_data, im_info):
        """The engine's two halves as CUDA graphs (the geometry, every buffer and every plan are fixed): the forward, and
        the backward in one graph per gradient bucket, so that the bucket's all-reduce can be issued between two replays.
        Removes ~900 kernel launches of host work per step; the heads in between stay eager (data-dependent shapes)."""
        self.im_static, self.info_static = im_data.clone(), im_info.clone()
        self.grads_static = [torch.zeros_like(t) for t in (self.cls_map, self.bbox_map, self.rpn_score, self.rpn_delta,
                                                            self.trk_layer.out_nchw)]
        torch.cuda.synchronize()
        count = ops.LAUNCHES
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd):
            self.info_graph = self._engine_forward(self.im_static, self.info_static)
        self._fwd_launches = ops.LAUNCHES - count
        pool = self.g_fwd.pool()
        self.g_bwd = []
        for a, b in self._segments():
            count = ops.LAUNCHES
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                if a == 0:
                    self._load_leaf_grads(self.grads_static)
                for k in range(a, b):
                    self.bwd[k][0]()
            self.g_bwd.append((g, ops.LAUNCHES - count))
        if self.prepack:
            count = ops.LAUNCHES
            self.g_xpack = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_xpack, pool=pool):
                self._prepack_inputs()
            self._xpack_launches = ops.LAUNCHES - count
        count = ops.LAUNCHES
        self.g_refresh = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_refresh, pool=pool):
            self._refresh_weights()
        self._refresh_launches = ops.LAUNCHES - count

    def _heads(self, info, gt_boxes, num_boxes):
        """proposal step, target layers, PSRoI heads and the five losses on autograd LEAVES of the five convolution outputs
        they consume, then d(loss)/d(leaves) (trainval_net.py:367-368).  No device->host round trip anywhere in here, so
        the whole thing -- forward and autograd backward -- replays as one CUDA graph."""
        from model.rpn.proposal_target_layer_cascade import train_heads
        leaves = [t.detach().requires_grad_() for t in (self.cls_map, self.bbox_map, self.rpn_score, self.rpn_delta,
                                                        self.trk_layer.out_nchw)]
        cls_map, bbox_map, score, delta, trk = leaves
        out = train_heads(self.net, self.B, None, None, None, None, cls_map, bbox_map, info, gt_boxes, num_boxes,
                          rpn_maps=(score, delta), trk_map=trk)
        loss = out[4].mean() + out[5].mean() + out[6].mean() + out[7].mean() + out[9].mean()   # trainval_net.py:367-368
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        grads = [gr if gr is not None else torch.zeros_like(leaf) for gr, leaf in zip(grads, leaves)]
        return out, loss, grads

    def _capture_heads(self, gt_boxes, num_boxes):
        """the heads as a CUDA graph of their own (private memory pool: it is captured after, but replayed between, the
        forward and backward graphs)"""
        self.gt_static, self.nb_static = gt_boxes.clone(), num_boxes.clone()
        torch.cuda.synchronize()
        count = ops.LAUNCHES
        g = torch.cuda.CUDAGraph()
        # captured on a HIGH-priority stream: the graph's kernel nodes keep that priority, so the heads' many small kernels
        # are scheduled ahead of the pending blocks of the input-plane packers that run beside them (pack_stream)
        hp = torch.cuda.Stream(device=self.gt_static.device, priority=-1) if os.environ.get("D2T_HEADS_PRIORITY", "1") != "0" else None
        with torch.cuda.graph(g, stream=hp):
            out, loss, grads = self._heads(self.info_graph, self.gt_static, self.nb_static)
            with torch.no_grad():
                for dst, gr in zip(self.grads_static, grads):
                    dst.copy_(gr)
        self._heads_static = (tuple(t.detach() if torch.is_tensor(t) else t for t in out), loss.detach())
        self._heads_launches = ops.LAUNCHES - count
        self.g_heads = g

    def forward_backward(self, im_data, im_info, gt_boxes, num_boxes):
        """forward (training mode) + losses + backward; fills param.grad (mean over ranks when torch.distributed is
        initialised).  Returns (the reference's 10-tuple, total loss); with CUDA graphs (the default, from the third call
        on) these are the graphs' static output tensors, overwritten by the next call."""
        net, B = self.net, self.B
        assert net.training, "call net.train() first (the RPN picks its TRAIN configuration from it)"
        self._calls = getattr(self, "_calls", 0) + 1
        if self.use_graphs and self.g_fwd is None and self._calls > 2:       # (two eager steps first: every kernel variant
            with torch.no_grad():                                            # has been launched once before the capture)
                self._capture(im_data, im_info)
            if self.graph_heads:
                self._capture_heads(gt_boxes, num_boxes)
        graphed = self.g_fwd is not None
        with torch.no_grad():
            if graphed:
                self.im_static.copy_(im_data, non_blocking=True)
                self.info_static.copy_(im_info, non_blocking=True)
                self.g_fwd.replay()
                ops._count(self._fwd_launches)
                info = self.info_graph
            else:
                info = self._engine_forward(im_data, im_info)
            if self.prepack:           # the weight gradients' input planes: beside the heads (side stream) once graphed
                if graphed and self.g_xpack is not None:
                    self.pack_stream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(self.pack_stream):
                        self.g_xpack.replay()
                    ops._count(self._xpack_launches)
                else:
                    self._prepack_inputs()
        if graphed and self.g_heads is not None:
            with torch.no_grad():
                self.gt_static.copy_(gt_boxes, non_blocking=True)
                self.nb_static.copy_(num_boxes, non_blocking=True)
            self.g_heads.replay()                                            # (fills grads_static)
            ops._count(self._heads_launches)
            out, loss = self._heads_static
            self.leaf_grads = self.grads_static
            with torch.no_grad():
                self._run_backward()
            return out, loss
        out, loss, self.leaf_grads = self._heads(info, gt_boxes, num_boxes)
        with torch.no_grad():
            if graphed:
                for dst, gr in zip(self.grads_static, self.leaf_grads):
                    dst.copy_(gr, non_blocking=True)
            else:
                self._load_leaf_grads([gr.contiguous() for gr in self.leaf_grads])
            self._run_backward()
        return out, loss

    def _run_backward(self):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        main = torch.cuda.current_stream()
        if self.prepack:
            main.wait_stream(self.pack_stream)            # the input planes are in place
        self._comm_events = []
        graphed = self.g_fwd is not None
        for i, (a, b) in enumerate(self._segments()):
            if graphed:
                g, n = self.g_bwd[i]
                g.replay()
                ops._count(n)
            else:
                for k in range(a, b):
                    self.bwd[k][0]()
            if world > 1 and i < len(self.buckets):
                _, lo, hi = self.buckets[i]
                self.ready.record(main)
                self.comm_stream.wait_event(self.ready)
                with torch.cuda.stream(self.comm_stream):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    part = self.flat[lo:hi]
                    e0.record(self.comm_stream)
                    if dist.get_backend() == "nccl":
                        dist.all_reduce(part, op=dist.ReduceOp.AVG)
                    else:
                        dist.all_reduce(part)
                        part.div_(world)
                    e1.record(self.comm_stream)
                    self._comm_events.append((e0, e1))
        if world > 1:
            self._main_done = torch.cuda.Event(enable_timing=True)
            self._main_done.record(main)
            main.wait_stream(self.comm_stream)

    def comm_stats(self):
        """after a synchronised step: (sum of the buckets' all-reduce durations, the part of it the backward pass did NOT
        hide = how long the main stream had to wait after its last backward kernel), both in ms"""
        if not getattr(self, "_comm_events", None):
            return 0.0, 0.0
        total = sum(a.elapsed_time(b) for a, b in self._comm_events)
        exposed = max(0.0, self._main_done.elapsed_time(self._comm_events[-1][1]))
        return total, exposed
