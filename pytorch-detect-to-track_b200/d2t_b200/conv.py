"""Tensor-level front end of the tcgen05 convolution engine (csrc/conv.cu, csrc/conv_util.cu).

``ActTensor`` is a plain fp32 NHWC activation buffer plus one device float, ``amax`` -- a running max |x| that the
producing kernel's epilogue maintains and the fp16-split conv mode (``passes=16``) turns into its power-of-two
operand scale; ``ConvLayer`` / ``StemConv`` / ``CorrLayer`` own packed weights (hi, lo), the folded BatchNorm / bias
vectors, their output buffers and the C-ABI plan (TMA tensor maps) that runs them.  Nothing here computes on the CPU;
everything is enqueued on torch's current stream.

``passes``: 3 = 3xTF32, 16 = 3xFP16 (both fp32-accurate), 1 = single-pass TF32.
"""
import ctypes as C
import math

import torch

from ._lib import ConvDesc, D2TError, check, lib
from . import ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pad32(c):
    return (c + 31) // 32 * 32


def _pad64(c):
    return (c + 63) // 64 * 64


class AmaxArena(object):
    """One device buffer for the amax scalars of every ActTensor created inside ``with arena:`` -- an engine zeroes
    them all with a single memset per forward instead of one per tensor."""
    _active = None

    def __init__(self, slots, device="cuda"):
        self.buf = torch.zeros(slots, device=device)
        self.used = 0

    def take(self):
        if self.used >= self.buf.numel():
            raise D2TError("AmaxArena exhausted")
        self.used += 1
        return self.buf[self.used - 1:self.used]

    def zero(self):
        self.buf.zero_()
        ops._count(1)

    def __enter__(self):
        self._prev, AmaxArena._active = AmaxArena._active, self
        return self

    def __exit__(self, *exc):
        AmaxArena._active = self._prev
        return False


def _p(t):
    return t.data_ptr() if t is not None else None


class ActTensor(object):
    """[N, H, W, cstride] fp32; channels [0, C) are meaningful, the rest are zero."""

    def __init__(self, N, H, W, C_, cstride=None, device="cuda", amax=None, zero=True):
        self.N, self.H, self.W, self.C = N, H, W, C_
        self.cstride = cstride if cstride is not None else (C_ + 3) // 4 * 4
        self.x = (torch.zeros if zero else torch.empty)(N, H, W, self.cstride, device=device)
        arena = AmaxArena._active
        self.arena_owned = amax is None and arena is not None      # an engine zeroes it once per forward
        self.amax = amax if amax is not None else (arena.take() if arena is not None else torch.zeros(1, device=device))

    @staticmethod
    def from_nchw(x, cstride=None, amax=True):
        N, C_, H, W = x.shape
        t = ActTensor(N, H, W, C_, cstride if cstride is not None else _pad32(C_), x.device)
        return t.load_nchw(x, amax=amax)

    def load_nchw_amax(self, x):
        """the whole buffer <- x [N, C, H, W] (padding channels zero) and amax <- max |x|, in ONE launch after a 4-byte memset"""
        N, C_, H, W = x.shape
        self.amax.zero_()
        check(lib().d2t_nchw_to_nhwc_amax(x.contiguous().data_ptr(), N, C_, H, W, self.cstride, 0, self.cstride,
                                          self.x.data_ptr(), self.amax.data_ptr(), _stream()), "d2t_nchw_to_nhwc_amax")
        ops._count(2)
        return self

    def load_nchw(self, x, coffset=0, cwidth=None, amax=True):
        """write x [N, C, H, W] into channels [coffset, coffset + cwidth) (x's channels, then zeros)"""
        N, C_, H, W = x.shape
        cwidth = cwidth if cwidth is not None else self.cstride - coffset
        check(lib().d2t_nchw_to_nhwc(x.contiguous().data_ptr(), N, C_, H, W, self.cstride, coffset, cwidth,
                                     self.x.data_ptr(), _stream()), "d2t_nchw_to_nhwc")
        ops._count(1)
        if amax:   # (not on the engine path; only a 3xFP16 consumer needs it)
            torch.maximum(self.amax, x.detach().abs().max().reshape(1).float(), out=self.amax)
        return self

    def batch_slice(self, n0, n1):
        """view of images [n0, n1) (contiguous in NHWC)"""
        v = ActTensor.__new__(ActTensor)
        v.N, v.H, v.W, v.C, v.cstride = n1 - n0, self.H, self.W, self.C, self.cstride
        v.x = self.x[n0:n1]
        v.amax, v.arena_owned = self.amax, self.arena_owned       # (a bound for the whole tensor bounds the slice)
        return v

    def to_nchw(self, C_=None, coffset=0):
        C_ = self.C if C_ is None else C_
        out = torch.empty(self.N, C_, self.H, self.W, device=self.x.device)
        check(lib().d2t_nhwc_to_nchw(self.x.data_ptr(), self.N, C_, self.H, self.W, self.cstride, coffset,
                                     out.data_ptr(), _stream()), "d2t_nhwc_to_nchw")
        ops._count(1)
        return out


def pack_weights(w, cin_pad=None, lo=True):
    """OIHW -> ([O, R*S*cin_pad] w, w_lo)"""
    O, I, R, S = w.shape
    cin_pad = cin_pad or _pad32(I)
    w = w.detach().contiguous().float()
    hi = torch.empty(O, R * S * cin_pad, device=w.device)
    lo_t = torch.empty_like(hi) if lo else None
    check(lib().d2t_conv_pack_weights(w.data_ptr(), O, I, R, S, cin_pad, hi.data_ptr(), _p(lo_t), _stream()),
          "d2t_conv_pack_weights")
    torch.cuda.current_stream().synchronize()   # `w` may be a temporary
    return hi, lo_t


def pack_weights_f16(w, cin_pad=None, scale=None):
    """OIHW -> ([O, R*S*cin_pad] fp16 hi, lo, w_exp): the pair splits w * scale[o] * 2^w_exp, max |.| * 2^w_exp in
    [2^14, 2^15).  `scale`: the convolution's per-output-channel factor (folded BatchNorm), carried by the weights."""
    O, I, R, S = w.shape
    cin_pad = cin_pad or _pad64(I)
    w = w.detach().contiguous().float()
    scale = scale.detach().float().contiguous() if scale is not None else None
    wmax = float((w * scale.view(-1, 1, 1, 1)).abs().max()) if scale is not None else float(w.abs().max())
    w_exp = 15 - math.frexp(wmax)[1] if wmax > 0 and math.isfinite(wmax) else 0
    w_exp = max(-100, min(100, w_exp))
    hi = torch.empty(O, R * S * cin_pad, device=w.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    check(lib().d2t_conv_pack_weights_f16(w.data_ptr(), _p(scale), O, I, R, S, cin_pad, w_exp, hi.data_ptr(), lo.data_ptr(),
                                          _stream()), "d2t_conv_pack_weights_f16")
    torch.cuda.current_stream().synchronize()   # `w` may be a temporary
    return hi, lo, w_exp


class _Planned(object):
    plan = None
    zero_amax = None      # the output's amax scalar when this layer owns zeroing it (stand-alone use)

    def set_scratch(self, scratch):
        """give this plan a private stream-K scratch (uint8 CUDA tensor of d2t_conv_scratch_bytes(), zero-initialised)
        so that its chain may run concurrently with other chains on another stream"""
        check(lib().d2t_conv_plan_set_scratch(self.plan, scratch.data_ptr(), scratch.numel()), "d2t_conv_plan_set_scratch")
        self._scratch = scratch

    def set_early_weights(self, on=True):
        """fetch the first weight tiles before the launch waits for its predecessor (d2t_conv_plan_set_early_weights):
        for layers whose packed weights are not rewritten by the launch immediately before them"""
        check(lib().d2t_conv_plan_set_early_weights(self.plan, int(on)), "d2t_conv_plan_set_early_weights")

    def set_done(self, prev, self_counter):
        """completion hand-shake with the layer launched immediately before this one on the same stream (csrc/conv.cu,
        d2t_conv_plan_set_done): `self_counter` is a one-element int32 CUDA tensor this layer's CTAs count into, `prev` the
        previous layer (whose counter this one polls instead of griddepcontrol.wait) or None.  The caller zeroes the
        counters before every pass over the chain."""
        self.done_counter = self_counter
        check(lib().d2t_conv_plan_set_done(self.plan, prev.plan if prev is not None else None,
                                           _p(prev.done_counter) if prev is not None else None, _p(self_counter)),
              "d2t_conv_plan_set_done")

    def _bind_amax(self, x, out):
        """attach the input's / output's amax scalars; outside an engine arena the layer zeroes its output's before
        each run (inside one, several producers may share an output buffer and the engine zeroes all at once)"""
        check(lib().d2t_conv_plan_set_amax(self.plan, _p(x.amax) if x is not None else None,
                                           _p(out.amax) if out is not None else None), "d2t_conv_plan_set_amax")
        if out is not None and not out.arena_owned:
            self.zero_amax = out.amax

    def run(self, stream=None):
        """enqueue on `stream` (a raw cudaStream_t handle) or, by default, on torch's current stream"""
        if self.zero_amax is not None:
            self.zero_amax.zero_()
        check(lib().d2t_conv_plan_run(self.plan, _stream() if stream is None else stream), "d2t_conv_plan_run")
        ops._count(1)

    def __del__(self):
        try:
            if self.plan:
                lib().d2t_conv_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass


class ConvLayer(_Planned):
    """out = relu?(scale * conv(x, w) + shift + residual) bound to fixed input / output buffers."""

    def __init__(self, x, weight, scale=None, shift=None, stride=1, pad=0, dil=1, relu=False, residual=None,
                 passes=3, out=None, out_coffset=0, want_nhwc=True, want_nchw=False, out_nchw=None, amax_w=None,
                 packed=None, mask=None):
        """passes = 16: `scale` (the folded BatchNorm factor) is multiplied into the packed weights and `shift` initialises
        the accumulator -- the kernel's epilogue has no affine pass.
        amax_w: one-element CUDA float tensor >= max |weight * scale| -- the weights are then (re)packed on the device from
        it (``repack()``, no host round trip: a training loop calls it after every optimizer step) and the kernel reads the
        scale from the same scalar.  packed = (hi, lo): operand matrices packed by the caller (backward-data, with amax_w);
        `weight` then only gives the geometry [O, I, R, S].  mask: ActTensor shaped like the output; out = mask > 0 ? . : 0."""
        O, I, R, S = weight.shape
        if _pad32(I) > x.cstride:
            raise ValueError("input buffer has %d channels per pixel, conv needs %d" % (x.cstride, _pad32(I)))
        self.x, self.residual, self.mask = x, residual, mask
        self.weight, self.amax_w = weight, amax_w
        dev = x.x.device
        self.scale = scale.detach().float().contiguous().to(dev) if scale is not None else None
        self.shift = shift.detach().float().contiguous().to(dev) if shift is not None else None
        w_exp = 0
        if packed is not None:
            assert passes == 16 and amax_w is not None
            self.w_hi, self.w_lo = packed
        elif amax_w is not None:
            assert passes == 16, "device-side weight scales are a 3xFP16 feature"
            self.w_hi = torch.empty(O, R * S * _pad64(I), device=weight.device, dtype=torch.float16)
            self.w_lo = torch.empty_like(self.w_hi)
            self.repack()
        elif passes == 16:
            self.w_hi, self.w_lo, w_exp = pack_weights_f16(weight, _pad64(I), self.scale)
        else:
            self.w_hi, self.w_lo = pack_weights(weight, _pad32(I), lo=(passes == 3))
        plan_scale = None if passes == 16 else self.scale          # 3xFP16: the scale is in the packed weights
        OH = (x.H + 2 * pad - dil * (R - 1) - 1) // stride + 1
        OW = (x.W + 2 * pad - dil * (S - 1) - 1) // stride + 1
        self.out = out if out is not None else (ActTensor(x.N, OH, OW, O, device=dev) if want_nhwc else None)
        if out_nchw is not None:
            assert tuple(out_nchw.shape) == (x.N, O, OH, OW) and out_nchw.is_contiguous()
        self.out_nchw = out_nchw if out_nchw is not None else (torch.empty(x.N, O, OH, OW, device=dev) if want_nchw else None)
        d = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=_pad32(I), in_cstride=x.cstride, Cout=O, R=R, S=S, stride=stride, pad=pad,
                     dil=dil, passes=passes, relu=int(relu), out_cstride=self.out.cstride if self.out is not None else 0,
                     out_coffset=out_coffset, res_cstride=residual.cstride if residual is not None else 0, w_exp=w_exp)
        self.plan = lib().d2t_conv_plan_create(
            C.byref(d), _p(x.x), _p(self.w_hi), _p(self.w_lo), _p(plan_scale), _p(self.shift),
            _p(residual.x) if residual is not None else None, _p(self.out.x) if self.out is not None else None,
            _p(self.out_nchw))
        if not self.plan:
            raise D2TError("d2t_conv_plan_create failed: %s" % lib().d2t_last_error().decode())
        info = (C.c_int * 8)()
        lib().d2t_conv_plan_info(self.plan, info)
        self.info = dict(zip(("OH", "OW", "tile_h", "tile_w", "BN", "m_tiles", "n_tiles", "grid"), list(info)))
        self.flops = 2.0 * x.N * OH * OW * O * I * R * S
        self._bind_amax(x, self.out)
        if amax_w is not None:
            check(lib().d2t_conv_plan_set_weight_amax(self.plan, _p(amax_w)), "d2t_conv_plan_set_weight_amax")
        if mask is not None:
            assert (mask.N, mask.H, mask.W) == (x.N, OH, OW) and mask.cstride >= O
            check(lib().d2t_conv_plan_set_mask(self.plan, _p(mask.x), mask.cstride), "d2t_conv_plan_set_mask")

    def repack(self):
        """re-derive the packed fp16 operand pair from the current values of `weight` (scale from *amax_w)"""
        O, I, R, S = self.weight.shape
        w = self.weight.detach()
        assert w.is_contiguous() and w.dtype == torch.float32
        check(lib().d2t_conv_pack_weights_f16_dev(w.data_ptr(), _p(self.scale), O, I, R, S, _pad64(I), _p(self.amax_w),
                                                  _p(self.w_hi), _p(self.w_lo), _stream()), "d2t_conv_pack_weights_f16_dev")
        ops._count(1)

    def repack_item(self):
        """arguments of repack() as a record for RepackMany"""
        O, I, R, S = self.weight.shape
        return (self.weight.detach(), self.scale, self.amax_w, self.w_hi, self.w_lo, O, I, R, S, _pad64(I), 0, 0)

    def run(self, stream=None):
        _Planned.run(self, stream)
        return self.out if self.out is not None else self.out_nchw


class RepackMany(object):
    """every packed operand of an engine refreshed by ONE launch (csrc/conv_util.cu: repack_many) instead of one small
    kernel per operand; `layers`: ConvLayer / DgradConv objects built with a device-side weight scale (amax_w)"""

    def __init__(self, layers):
        import struct
        assert lib().d2t_conv_repack_item_bytes() == 72
        recs, first, owner = [], 0, []
        self.keep = []
        for l in layers:
            w, scale, amax, hi, lo, O, I, R, S, pad, rows, dgrad = l.repack_item()
            assert w.is_contiguous() and w.dtype == torch.float32
            self.keep.append((w, scale, amax, hi, lo))
            nblk = lib().d2t_conv_repack_item_blocks(O, I, R, S, pad, rows, dgrad)
            assert nblk > 0, "repack_many: unsupported filter size"
            recs.append(struct.pack("<QQQQQiiiiiiii", w.data_ptr(), scale.data_ptr() if scale is not None else 0, amax.data_ptr(),
                                    hi.data_ptr(), lo.data_ptr(), O, I, R, S, pad, rows, dgrad, first))
            owner.append(torch.full((nblk,), len(recs) - 1, dtype=torch.int32))
            first += nblk
        self.n, self.blocks = len(recs), first
        self.items = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(hi.device)
        self.block_item = torch.cat(owner).to(hi.device)          # the item of every block

    def run(self, stream=None):
        check(lib().d2t_conv_repack_many(self.items.data_ptr(), self.n, self.blocks, self.block_item.data_ptr(),
                                         _stream() if stream is None else stream),
              "d2t_conv_repack_many")
        ops._count(1)


class ConvChain(object):
    """Consecutive ConvLayers run by ONE persistent launch (csrc/conv.cu: conv_chain; include/d2t_b200.h "Layer chains").
    The layers' plans are copied at construction -- build the chain after every set_scratch / set_done / amax binding.
    A grid-wide barrier separates a layer from the ones before it only where it reads (input or residual) a buffer that a
    layer since the previous barrier writes; results are bit-identical to ``for l in layers: l.run()``."""

    def __init__(self, layers):
        assert len(layers) > 0 and all(ConvChain.chainable(l) for l in layers)
        self.layers = list(layers)
        n = len(layers)

        def store(t):
            return t.untyped_storage().data_ptr() if t is not None else None

        written, sync = set(), []
        for l in layers:
            reads = {store(l.x.x), store(l.residual.x) if l.residual is not None else None} - {None}
            need = bool(reads & written)
            if need:
                written = set()
            sync.append(int(need))
            written |= {store(l.out.x) if l.out is not None else None, store(l.out_nchw)} - {None}
        sync[0] = 0
        self.sync_before = sync
        dev = layers[0].x.x.device
        self.buf = torch.zeros(lib().d2t_conv_chain_bytes(n) + 256, dtype=torch.uint8, device=dev)
        off = (-self.buf.data_ptr()) % 256
        plans = (C.c_void_p * n)(*[l.plan for l in layers])
        syncs = (C.c_int * n)(*sync)
        self.chain = lib().d2t_conv_chain_create(plans, syncs, n, self.buf.data_ptr() + off, self.buf.numel() - off)
        if not self.chain:
            raise D2TError("d2t_conv_chain_create failed: %s" % lib().d2t_last_error().decode())
        self.flops = sum(l.flops for l in layers)

    @staticmethod
    def chainable(layer):
        return isinstance(layer, ConvLayer) and layer.zero_amax is None and bool(lib().d2t_conv_plan_chainable(layer.plan))

    def run(self, stream=None):
        check(lib().d2t_conv_chain_run(self.chain, _stream() if stream is None else stream), "d2t_conv_chain_run")
        ops._count(1)

    def __del__(self):
        try:
            if self.chain:
                lib().d2t_conv_chain_destroy(self.chain)
                self.chain = None
        except Exception:
            pass


def build_chains(layers, min_len=2):
    """run list for `layers`: maximal runs of chainable ConvLayers become ConvChains, everything else stays as it is"""
    out, run = [], []

    def flush():
        if len(run) >= min_len:
            out.append(ConvChain(run))
        else:
            out.extend(run)
        del run[:]

    for l in layers:
        if ConvChain.chainable(l):
            run.append(l)
        else:
            flush()
            out.append(l)
    flush()
    return out


class DgradConv(ConvLayer):
    """Backward-data of  y = scale * conv(x, weight)  (stride 1: a strided 1x1 convolution is run at the output
    resolution and scattered by ``upsample2_add_mask``):  out = mask > 0 ? conv(g, wt) + residual : 0  with
    wt[ci][r'][s'][co] = weight[co][ci][R-1-r'][S-1-s'] * scale[co], padding dil * (R - 1) - pad, on the same tcgen05
    kernel.  `amax_wt`: one-element CUDA float tensor >= max |wt| (the caller keeps it current); ``repack()`` after
    every optimizer step.  `out_channels` >= Cin widens the output with zero channels (a padded forward input)."""

    def __init__(self, g, weight, scale, pad, dil, amax_wt, out=None, residual=None, mask=None, out_channels=None):
        O, I, R, S = weight.shape
        rows = out_channels if out_channels is not None else I
        self.fwd_weight, self.fwd_scale, self.rows = weight, (scale.detach().float().contiguous() if scale is not None else None), rows
        self.amax_w = amax_wt
        hi = torch.empty(rows, R * S * _pad64(O), device=weight.device, dtype=torch.float16)
        lo = torch.empty_like(hi)
        self.w_hi, self.w_lo = hi, lo
        self.repack()
        geom = torch.empty(rows, O, R, S, device="meta")              # [O', I', R, S] of the equivalent forward conv
        ConvLayer.__init__(self, g, geom, None, None, 1, dil * (R - 1) - pad, dil, False, residual, passes=16, out=out,
                           amax_w=amax_wt, packed=(hi, lo), mask=mask)

    def repack_item(self):
        O, I, R, S = self.fwd_weight.shape
        return (self.fwd_weight.detach(), self.fwd_scale, self.amax_w, self.w_hi, self.w_lo, O, I, R, S, _pad64(O), self.rows, 1)

    def repack(self):
        O, I, R, S = self.fwd_weight.shape
        w = self.fwd_weight.detach()
        check(lib().d2t_conv_pack_weights_f16_dgrad(w.data_ptr(), _p(self.fwd_scale), O, I, self.rows, R, S, _pad64(O),
                                                    _p(self.amax_w), _p(self.w_hi), _p(self.w_lo), _stream()),
              "d2t_conv_pack_weights_f16_dgrad")
        ops._count(1)


class WgradScratch(object):
    """The weight-gradient GEMM reads both operands from channel-major planes (csrc/conv.cu, WGRAD): one reusable pair of
    plane buffers, sized for the largest layer, shared by every WgradLayer of an engine (plans run one after another)."""

    def __init__(self, x_floats, g_halfs, device="cuda"):
        self.xt = torch.zeros(x_floats, device=device)
        self.g_hi = torch.zeros(g_halfs, device=device, dtype=torch.float16)
        self.g_lo = torch.zeros(g_halfs, device=device, dtype=torch.float16)
        # partial tiles of the split-K reduction (d2t_wgrad_plan_set_partials); D2T_WGRAD_REDUCE=0: in-kernel finisher
        import os
        self.partials = (torch.empty(lib().d2t_wgrad_partials_bytes(), dtype=torch.uint8, device=device)
                         if os.environ.get("D2T_WGRAD_REDUCE", "1") != "0" else None)

    @staticmethod
    def need(x, g, stride, S=1):
        """(floats of input planes, halfs of gradient planes: one column-shifted copy per filter column)"""
        return x.N * x.C * g.H * _pad32(g.W), S * g.N * g.C * g.H * _pad32(g.W)


class WgradLayer(_Planned):
    """grad_w [O, I, R, S] = scale[o] * sum_pixels x (*) g  for  y = scale * conv(x, w, stride, pad, dil):  x the forward
    input (ActTensor), g the gradient w.r.t. the convolution's pre-activation output (ActTensor, amax current).  Three
    launches: plane packs of x and g, then the tcgen05 GEMM, which writes every element of `grad_w`."""

    def __init__(self, x, g, grad_w, scale, stride, pad, dil, scratch, cin=None, own_xt=False):
        O, I, R, S = grad_w.shape
        assert grad_w.is_contiguous() and grad_w.dtype == torch.float32
        assert stride == 1 or (R == 1 and S == 1 and pad == 0), "strided convolutions: 1x1 only"
        self.x, self.g, self.grad_w, self.stride = x, g, grad_w, stride
        self.scale = scale.detach().float().contiguous() if scale is not None else None
        self.xh, self.xw = (g.H, g.W) if stride > 1 else (x.H, x.W)
        self.xp, self.gp = _pad32(self.xw), _pad32(g.W)
        self.scratch = scratch
        self.S, self.pad, self.dil = S, pad, dil
        nx, ng = x.N * I * self.xh * self.xp, S * g.N * O * g.H * self.gp
        # own_xt: the layer keeps its OWN input planes, packed ahead of the backward pass by pack_input() (the forward
        # activations exist as soon as the forward is done: a training engine packs them beside the latency-bound heads)
        self.xt = torch.zeros(nx, device=grad_w.device) if own_xt else scratch.xt
        self.prepacked = bool(own_xt)
        if nx > self.xt.numel() or ng > scratch.g_hi.numel():
            raise ValueError("WgradScratch too small: need %d floats / %d halfs" % (nx, ng))
        self.plan = lib().d2t_wgrad_plan_create(x.N, I, O, self.xh, self.xw, self.xp, g.H, g.W, self.gp, R, S, pad, dil,
                                                _p(self.xt), _p(scratch.g_hi), _p(scratch.g_lo), _p(x.amax), _p(g.amax),
                                                _p(self.scale), _p(grad_w))
        if not self.plan:
            raise D2TError("d2t_wgrad_plan_create failed: %s" % lib().d2t_last_error().decode())
        self.flops = 2.0 * g.N * g.H * g.W * O * I * R * S
        self.launches = 3
        if getattr(scratch, "partials", None) is not None:
            check(lib().d2t_wgrad_plan_set_partials(self.plan, scratch.partials.data_ptr(), scratch.partials.numel()),
                  "d2t_wgrad_plan_set_partials")
            self.launches = 4

    def pack_input(self, stream=None):
        """the forward input as channel-major planes (the A operand of the GEMM)"""
        x = self.x
        I = self.grad_w.shape[1]
        check(lib().d2t_wgrad_pack_input(_p(x.x), x.N, x.H, x.W, x.cstride, I, self.stride, self.xh, self.xw, self.xp,
                                         _p(self.xt), _stream() if stream is None else stream), "d2t_wgrad_pack_input")
        ops._count(1)

    def run(self, stream=None):
        x, g, sc = self.x, self.g, self.scratch
        O, I = self.grad_w.shape[:2]
        st = _stream() if stream is None else stream
        if not self.prepacked:
            self.pack_input(st)
        check(lib().d2t_wgrad_pack_grad(_p(g.x), g.N, g.H, g.W, g.cstride, O, self.gp, self.S, self.dil, self.pad,
                                        _p(g.amax), _p(sc.g_hi), _p(sc.g_lo), st), "d2t_wgrad_pack_grad")
        check(lib().d2t_conv_plan_run(self.plan, st), "d2t_conv_plan_run")
        ops._count(self.launches - 1)
        return self.grad_w


class CorrBwdScratch(object):
    """operand buffers of the correlation-backward GEMM (csrc/conv.cu, CORRB), reused by every CorrBwdLayer of an engine:
    the expanded band [N][H][W][D][64] and the other frame's planes [N][C][H][pitch], fp16 (hi, lo) each"""

    def __init__(self, band_halfs, other_halfs, device="cuda"):
        self.e_hi = torch.zeros(band_halfs, device=device, dtype=torch.float16)
        self.e_lo = torch.zeros(band_halfs, device=device, dtype=torch.float16)
        self.o_hi = torch.zeros(other_halfs, device=device, dtype=torch.float16)
        self.o_lo = torch.zeros(other_halfs, device=device, dtype=torch.float16)

    @staticmethod
    def need(N, C, H, W, r):
        return N * H * W * (2 * r + 1) * 64, N * C * H * ((W + 7) // 8 * 8)


class CorrBwdLayer(_Planned):
    """Gradient of the cross-frame correlation (kernel_size 1, stride1 == stride2 = stride, pad == max_displacement) with
    respect to ONE of its inputs, on the tensor cores:  which = 1: d/d(input1) from (gO, input2); which = 2: d/d(input2) from
    (gO, input1).  `gout`: NHWC buffer holding gO in channels [coff, coff + D*D) (its amax current); `other`: the other
    frame's NHWC features; `out`: NHWC gradient on the correlation LATTICE ([N, ceil(H/stride), ceil(W/stride), C]: for
    stride 2 the gradient lives on the even positions only) -- every element written.  Three launches."""

    def __init__(self, gout, coff, other, out, md, stride, which, scratch):
        r = md // stride
        self.r, self.D, self.which, self.stride = r, 2 * r + 1, which, stride
        self.gout, self.coff, self.other, self.out, self.scratch = gout, coff, other, out, scratch
        N, H, W, Cc = out.N, out.H, out.W, other.C
        assert (gout.N, gout.H, gout.W) == (N, H, W) and other.N == N
        assert H == -(-other.H // stride) and W == -(-other.W // stride)
        self.pitch = (W + 7) // 8 * 8
        nb, no = CorrBwdScratch.need(N, Cc, H, W, r)
        if nb > scratch.e_hi.numel() or no > scratch.o_hi.numel():
            raise ValueError("CorrBwdScratch too small")
        self.inv = torch.full((Cc,), 1.0 / Cc, device=out.x.device)
        self.plan = lib().d2t_corrb_plan_create(N, Cc, H, W, r, _p(scratch.e_hi), _p(scratch.e_lo), _p(scratch.o_hi),
                                                _p(scratch.o_lo), self.pitch, _p(gout.amax), _p(other.amax), _p(self.inv),
                                                _p(out.x), out.cstride, 0)
        if not self.plan:
            raise D2TError("d2t_corrb_plan_create failed: %s" % lib().d2t_last_error().decode())
        self.flops = 2.0 * N * H * W * self.D * self.D * Cc

    def run(self, stream=None):
        g, o, sc, out = self.gout, self.other, self.scratch, self.out
        st = _stream() if stream is None else stream
        check(lib().d2t_corrb_pack_band(_p(g.x), g.N, g.H, g.W, g.cstride, self.coff, self.r, int(self.which == 2), _p(g.amax),
                                        _p(sc.e_hi), _p(sc.e_lo), st), "d2t_corrb_pack_band")
        check(lib().d2t_corrb_pack_other(_p(o.x), o.N, o.H, o.W, o.cstride, o.C, self.stride, out.H, out.W, self.pitch,
                                         _p(o.amax), _p(sc.o_hi), _p(sc.o_lo), st), "d2t_corrb_pack_other")
        check(lib().d2t_conv_plan_run(self.plan, st), "d2t_conv_plan_run")
        ops._count(3)
        return out


def upsample2_add_mask(low, out, extra=None, mask=None):
    """out = mask > 0 ? (even positions: low) + extra : 0; max |out| -> out.amax (csrc/conv_util.cu)"""
    assert out.cstride == out.C == low.cstride and (extra is None or extra.cstride == out.cstride)
    assert mask is None or mask.cstride == out.cstride
    check(lib().d2t_upsample2_add_mask(_p(low.x), low.H, low.W, _p(extra.x) if extra is not None else None,
                                       _p(mask.x) if mask is not None else None, out.N, out.H, out.W, out.C, _p(out.x),
                                       _p(out.amax), _stream()), "d2t_upsample2_add_mask")
    ops._count(1)
    return out


class StemConv(_Planned):
    """conv1 7x7 / stride 2 / pad 3 (+ folded bn1 + ReLU) on the tcgen05 kernel via the row-window TMA map
    (csrc/conv.cu: d2t_conv_stem_plan_create).  passes = 16 (3xFP16): a K block is TWO filter rows, the weights are packed
    once as fp16 (hi, lo) with the BatchNorm scale folded in, and the image's max |x| (its operand scale) comes out of the
    input packing kernel."""

    def __init__(self, N, H, W, weight, scale, shift, relu=True, passes=3, device="cuda"):
        O, I, R, S = weight.shape
        assert (R, S) == (7, 7) and I <= 4
        self.N, self.C, self.H, self.W, self.passes = N, I, H, W, passes
        Hp, Wp = (H + 7) & ~1, W + 8
        self.packed = torch.empty(N, Hp, Wp, 4, device=device)
        w32 = torch.empty(O, 7 * 32, device=device)
        w32_lo = torch.empty(O, 7 * 32, device=device)
        w = weight.detach().float().contiguous()
        check(lib().d2t_stem_pack_weights(w.data_ptr(), O, I, w32.data_ptr(), w32_lo.data_ptr(), _stream()),
              "d2t_stem_pack_weights")
        torch.cuda.current_stream().synchronize()   # `w` may be a temporary
        self.scale = scale.detach().float().contiguous().to(device) if scale is not None else None
        self.shift = shift.detach().float().contiguous().to(device) if shift is not None else None
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        self.out = ActTensor(N, OH, OW, O, device=device)
        if passes == 16:
            ws = F_pad_rows(w32 * (self.scale.view(-1, 1) if self.scale is not None else 1.0))      # [O][8 rows x 32] = [O][4][64]
            wmax = float(ws.abs().max())
            w_exp = max(-100, min(100, 15 - math.frexp(wmax)[1])) if wmax > 0 and math.isfinite(wmax) else 0
            ws = ws * (2.0 ** w_exp)
            self.w_hi = ws.half().contiguous()                                     # (hi, lo) as the weight packers split: rn, rn
            self.w_lo = (ws - self.w_hi.float()).half().contiguous()
            self.amax_w = torch.full((1,), wmax, device=device)                    # the kernel derives the same 2^k from it
            arena = AmaxArena._active
            self.amax_in = arena.take() if arena is not None else torch.zeros(1, device=device)
            self._own_amax_in = arena is None
            plan_scale = None
        else:
            self.w_hi, self.w_lo, plan_scale = w32, w32_lo, self.scale
        self.plan = lib().d2t_conv_stem_plan_create(N, H, W, O, passes, _p(self.packed), _p(self.w_hi), _p(self.w_lo),
                                                    _p(plan_scale), _p(self.shift), int(relu), _p(self.out.x),
                                                    self.out.cstride)
        if not self.plan:
            raise D2TError("d2t_conv_stem_plan_create failed: %s" % lib().d2t_last_error().decode())
        self.flops = 2.0 * N * OH * OW * O * I * 49
        if passes == 16:
            check(lib().d2t_conv_plan_set_amax(self.plan, _p(self.amax_in), _p(self.out.amax)), "d2t_conv_plan_set_amax")
            check(lib().d2t_conv_plan_set_weight_amax(self.plan, _p(self.amax_w)), "d2t_conv_plan_set_weight_amax")
            if not self.out.arena_owned:
                self.zero_amax = self.out.amax
        else:
            self._bind_amax(None, self.out)

    def run(self, x, pairs=0):
        """x: [N, C, H, W] fp32 image batch; pairs > 0 (3xFP16 stem): x is the reference's [pairs, 2, C, H, W] frame-pair
        batch and the packed batch is leg-major (frame n = leg * pairs + pair) -- no permuted copy of the input"""
        if self.passes == 16:
            if self._own_amax_in:
                self.amax_in.zero_()
            check(lib().d2t_stem_pack_input_amax(x.data_ptr(), self.N, self.C, self.H, self.W, self.packed.data_ptr(),
                                                 self.amax_in.data_ptr(), int(pairs), _stream()), "d2t_stem_pack_input_amax")
        else:
            check(lib().d2t_stem_pack_input(x.data_ptr(), self.N, self.C, self.H, self.W, self.packed.data_ptr(), _stream()),
                  "d2t_stem_pack_input")
        ops._count(1)
        _Planned.run(self)
        return self.out


def F_pad_rows(w32):
    """[O][7 rows x 32] -> [O][8 rows x 32] (a zero eighth filter row: 3xFP16 K blocks are pairs of rows)"""
    return torch.cat([w32, torch.zeros(w32.size(0), 32, device=w32.device)], 1).contiguous()


class CorrLayer(_Planned):
    """Cross-frame correlation (kernel_size 1, stride1 == stride2) of two NHWC tensors on the tensor cores;
    writes a channel slice of `out` (NHWC) and/or a plain NCHW tensor.  passes = 16: 3xFP16 (kind::f16, twice the TF32
    rate; both inputs' amax scalars must be current), 3: 3xTF32, 1: single-pass TF32."""

    def __init__(self, x1, x2, pad, md, stride, passes=3, out=None, out_coffset=0, want_nchw=False):
        assert (x1.N, x1.H, x1.W, x1.cstride) == (x2.N, x2.H, x2.W, x2.cstride)
        r = md // stride
        self.D = 2 * r + 1
        oh = -(-(x1.H + 2 * pad - 2 * md) // stride)
        ow = -(-(x1.W + 2 * pad - 2 * md) // stride)
        self.x1, self.x2, self.out = x1, x2, out
        self.out_nchw = torch.empty(x1.N, self.D * self.D, oh, ow, device=x1.x.device) if want_nchw else None
        self.plan = lib().d2t_corr_plan_create(x1.N, _pad32(x1.C), x1.C, x1.H, x1.W, x1.cstride, pad, md, stride, passes,
                                               _p(x1.x), _p(x2.x), _p(out.x) if out is not None else None,
                                               out.cstride if out is not None else 0, out_coffset, _p(self.out_nchw))
        if not self.plan:
            raise D2TError("d2t_corr_plan_create failed: %s" % lib().d2t_last_error().decode())
        self.flops = 2.0 * x1.N * oh * ow * self.D * self.D * x1.C
        self._bind_amax(x1 if passes == 16 else None, out)
        if passes == 16:      # 3xFP16: both operands are activations, each scaled by its own tensor's amax
            check(lib().d2t_conv_plan_set_weight_amax(self.plan, _p(x2.amax)), "d2t_conv_plan_set_weight_amax")

    def run(self, stream=None):
        _Planned.run(self, stream)
        return self.out_nchw if self.out_nchw is not None else self.out


def maxpool3x3s2(x, out=None):
    OH = -(-(x.H - 3) // 2) + 1
    OW = -(-(x.W - 3) // 2) + 1
    if (OH - 1) * 2 >= x.H:
        OH -= 1
    if (OW - 1) * 2 >= x.W:
        OW -= 1
    # (the pooled tensor shares its input's amax: max-pooling cannot raise max |x|)
    out = out if out is not None else ActTensor(x.N, OH, OW, x.C, x.cstride, x.x.device, amax=x.amax)
    check(lib().d2t_maxpool3x3s2_nhwc(x.x.data_ptr(), x.N, x.H, x.W, x.cstride, out.x.data_ptr(), _stream()),
          "d2t_maxpool3x3s2_nhwc")
    ops._count(1)
    return out
