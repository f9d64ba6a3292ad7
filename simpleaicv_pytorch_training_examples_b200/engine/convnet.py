"""Forward/backward runtime of the ConvBnAct residual networks (ResNet, ResNetCifar).

The nn.Module shells in ``classification/backbones`` keep the reference's parameters, names and
constructors; this runtime executes their math on the B200 kernels of libsaicv_b200.so:

  conv     tcgen05 implicit GEMM (TMA im2col -> smem -> UMMA -> TMEM)      csrc/gemm_sm100.cuh
  BN       batch statistics / finalize / apply(+residual)(+ReLU), backward csrc/capi_elementwise.cu
  pooling  max 3x3/2, global average                                       csrc/capi_elementwise.cu

Activations are NHWC bf16; BatchNorm statistics, running stats, parameters and parameter
gradients are fp32 (the reference under autocast: SURVEY.md Appendix C).  Parameter gradients
are written straight into ``param.grad`` (or into the buffers a data-parallel wrapper handed
out through ``grad_buffer_of``), not returned through autograd.

Reference semantics followed: SimpleAICV/classification/backbones/resnet.py:19-48 (ConvBnActBlock),
:51-97 (BasicBlock), :100-155 (Bottleneck), :226-245 (ResNet.forward), resnetforcifar.py:98-108.
"""
import torch

from .. import ops

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


def _round_up(v, m):
    return (v + m - 1) // m * m


class GradSink:
    """Where parameter gradients go: ``param.grad`` by default, or bucket views supplied by the
    data-parallel wrapper (distributed.py) so that the all-reduce runs on flat buffers."""

    def __init__(self):
        self.views = {}       # id(param) -> fp32 tensor view with the parameter's shape
        self.on_ready = None  # callback(param) fired as soon as a gradient is complete
        self.on_backward_end = None  # callback() fired when the whole backward pass is enqueued

    def buffer_for(self, param):
        v = self.views.get(id(param))
        if v is None:
            v = torch.empty_like(param, dtype=torch.float32)
            self.views[id(param)] = v
        return v

    def begin(self, param):
        """Returns (buffer, accumulate): accumulate when param.grad already holds a gradient."""
        buf = self.buffer_for(param)
        if param.grad is None:
            return buf, False
        if param.grad.data_ptr() != buf.data_ptr():
            buf.copy_(param.grad)
        return buf, True

    def done(self, param, buf):
        param.grad = buf
        if self.on_ready is not None:
            self.on_ready(param)


class ConvBN:
    """Runtime state of one ConvBnActBlock: conv (no bias) -> BatchNorm2d -> activation."""

    def __init__(self, block, act):
        self.conv = block.layer[0]
        self.bn = block.layer[1]
        self.act = act
        k, c, r, s = self.conv.weight.shape
        self.k, self.c, self.r, self.s = k, c, r, s
        self.stride = self.conv.stride[0]
        self.pad = self.conv.padding[0]
        assert self.conv.groups == 1 and self.conv.bias is None and r == s
        assert self.conv.stride[0] == self.conv.stride[1] and self.conv.padding[0] == self.conv.padding[1]
        assert k % 8 == 0, 'output channels must be a multiple of 8'
        self.is_stem = (c % 64) != 0
        self.kpad = _round_up(r * s * c, 64) if self.is_stem else r * s * c
        self.w_bf16 = None
        self.w_version = None
        self.ss = None
        self.saved = None
        self.sums = None

    # ---- parameters
    def prep(self):
        w = self.conv.weight
        if self.w_bf16 is None or self.w_bf16.device != w.device:
            dev = w.device
            self.w_bf16 = torch.empty(self.k, self.kpad, device=dev, dtype=torch.bfloat16)
            self.ss = torch.empty(2, self.k, device=dev)
            self.saved = torch.empty(2, self.k, device=dev)
            self.sums = torch.zeros(2, self.k, device=dev)
            self.w_version = None
        ver = (w.data_ptr(), w._version)
        if ver != self.w_version:
            ops.prep_conv_weight(w.detach(), self.w_bf16, self.kpad,
                                 order=ops.ORDER_CRS if self.is_stem else ops.ORDER_RSC)
            self.w_version = ver

    # ---- forward
    def out_hw(self, h, w):
        return (ops.conv_out_size(h, self.pad, self.r, self.stride),
                ops.conv_out_size(w, self.pad, self.s, self.stride))

    def conv_fwd(self, a_in, tape):
        """a_in: NHWC bf16 activation, or the NCHW fp32 image batch for the stem."""
        if self.is_stem:
            n, _, h, w = a_in.shape
            P, Q = self.out_hw(h, w)
            cols = ops.stem_im2col(a_in, self.r, self.s, self.stride, self.pad, self.kpad)
            y = ops.linear_fwd(cols, self.w_bf16).view(n, P, Q, self.k)
            tape['cols'] = cols
        else:
            n, h, w, _ = a_in.shape
            cs = ops.make_conv_shape(n, h, w, self.c, self.k, self.r, self.s, self.stride, self.pad)
            y = ops.conv_fprop(a_in, self.w_bf16, cs)
            tape['a_in'] = a_in
            tape['cs'] = cs
        tape['in_hw'] = (h, w)
        tape['y'] = y
        return y

    def bn_prepare(self, y, training):
        """Fills self.ss (scale/shift) from batch statistics (training) or running stats (eval)."""
        bn = self.bn
        rows = y.numel() // self.k
        if training or not bn.track_running_stats:
            momentum = bn.momentum if bn.momentum is not None else 0.1
            partials = ops.bn_stats(y)
            ops.bn_finalize(partials, bn.weight.detach(), bn.bias.detach(),
                            bn.running_mean if bn.track_running_stats else None,
                            bn.running_var if bn.track_running_stats else None,
                            self.ss, self.saved, rows, bn.eps, momentum)
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
        else:
            scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
            self.ss[0].copy_(scale)
            self.ss[1].copy_(bn.bias.detach() - bn.running_mean * scale)

    def forward(self, a_in, tape, training, res=None, res_unit=None):
        """conv -> BN -> (+res) -> act.  res_unit: ConvBN whose BN is applied to `res` on the fly
        (downsample branch)."""
        y = self.conv_fwd(a_in, tape)
        self.bn_prepare(y, training)
        out = torch.empty_like(y)
        ops.bn_apply(y, self.ss, out, self.act, res=res,
                     res_scale_shift=res_unit.ss if res_unit is not None else None)
        tape['out'] = out
        return out

    # ---- backward
    def bn_bwd(self, dout, tape, sink, act_out=None, want_dres=False, act=None):
        """dout: gradient w.r.t. the activated output.  Returns (dy, dres).

        The ReLU mask comes from `act_out` (block output, when a residual was added before the
        activation) or, for plain conv->BN->ReLU units, is recomputed inside the kernels from
        sign(y*scale+shift), which saves reading the activated tensor twice."""
        act = self.act if act is None else act
        y = tape['y']
        mask_src = act_out if act != ACT_NONE else None
        ss = self.ss if (act != ACT_NONE and mask_src is None) else None
        ops.bn_bwd_reduce(dout, mask_src, y, self.saved, self.sums, act, scale_shift=ss)
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if want_dres else None
        gbuf, gacc = sink.begin(self.bn.weight)
        bbuf, bacc = sink.begin(self.bn.bias)
        assert gacc == bacc
        ops.bn_bwd_apply(dout, mask_src, y, self.saved, self.bn.weight.detach(), self.sums, dy, dres,
                         gbuf, bbuf, act, accumulate=gacc, scale_shift=ss)
        sink.done(self.bn.weight, gbuf)
        sink.done(self.bn.bias, bbuf)
        return dy, dres

    def conv_bwd(self, dy, tape, sink, need_dx=True, add=None):
        """dy: gradient w.r.t. the raw conv output [n,P,Q,k].  Returns the data gradient
        (NHWC bf16, `add` fused in when given) or None.  A 1x1 stride-2 conv returns
        ('strided', dd) with the compact gradient dd[n,P,Q,c] that belongs at the even pixels."""
        w = self.conv.weight
        wbuf, wacc = sink.begin(w)
        n, P, Q, _ = dy.shape
        h, wd = tape['in_hw']
        if self.is_stem:
            part = ops.linear_wgrad(dy.view(-1, self.k), tape['cols'])
            ops.finish_conv_wgrad(part, wbuf, self.kpad, accumulate=wacc, order=ops.ORDER_CRS)
            sink.done(w, wbuf)
            assert not need_dx, 'the stem has no data gradient'
            return None
        cs = tape['cs']
        part = ops.conv_wgrad(dy, tape['a_in'], cs)
        ops.finish_conv_wgrad(part, wbuf, self.kpad, accumulate=wacc)
        sink.done(w, wbuf)
        if not need_dx:
            return None
        if self.stride == 1:
            return ops.conv_dgrad(dy, self.w_bf16, cs, add=add)
        assert self.stride == 2
        if self.r == 1:
            assert add is None
            return 'strided', ops.linear_dgrad(dy.view(-1, self.k), self.w_bf16).view(n, P, Q, self.c)
        u = ops.zero_upsample2(dy, h, wd)
        cs1 = ops.make_conv_shape(n, h, wd, self.c, self.k, self.r, self.s, 1, self.pad)
        return ops.conv_dgrad(u, self.w_bf16, cs1, add=add)


class ResidualBlockRT:
    """BasicBlock (resnet.py:51-97) or Bottleneck (:100-155): conv chain + identity/downsample."""

    def __init__(self, block):
        names = ['conv1', 'conv2'] + (['conv3'] if hasattr(block, 'conv3') else [])
        self.units = []
        for nm in names:
            # the last unit has no activation of its own: its ReLU is applied after the shortcut add
            self.units.append(ConvBN(getattr(block, nm), ACT_RELU))
        self.down = ConvBN(block.downsample_conv, ACT_NONE) if block.downsample else None

    def all_units(self):
        return self.units + ([self.down] if self.down is not None else [])

    def forward(self, a_in, tape, training):
        tapes = tape.setdefault('u', [dict() for _ in self.units])
        x = a_in
        for u, t in zip(self.units[:-1], tapes[:-1]):
            x = u.forward(x, t, training)
        last, tl = self.units[-1], tapes[-1]
        if self.down is not None:
            td = tape.setdefault('d', dict())
            yd = self.down.conv_fwd(a_in, td)
            self.down.bn_prepare(yd, training)
            out = last.forward(x, tl, training, res=yd, res_unit=self.down)
        else:
            out = last.forward(x, tl, training, res=a_in)
        return out

    def backward(self, dout, tape, sink):
        tapes = tape['u']
        last, tl = self.units[-1], tapes[-1]
        out = tl['out']
        # g = dout * relu'(out) flows to both the last BN and the shortcut
        dy, g = last.bn_bwd(dout, tl, sink, act_out=out, want_dres=(self.down is None))
        shortcut, strided = g, None
        if self.down is not None:
            td = tape['d']
            dyd, _ = self.down.bn_bwd(dout, td, sink, act_out=out, act=ACT_RELU)
            shortcut = self.down.conv_bwd(dyd, td, sink)
            if isinstance(shortcut, tuple):
                strided, shortcut = shortcut[1], None
        dx = last.conv_bwd(dy, tl, sink)
        for u, t in zip(reversed(self.units[:-1]), reversed(tapes[:-1])):
            dy, _ = u.bn_bwd(dx, t, sink)
            dx = u.conv_bwd(dy, t, sink)
        # TODO(perf): fuse this add into the dgrad epilogue once the epilogue loads its residual
        # tile with TMA (the per-thread row-strided loads of EPI_RESID_BF16 are slower than this pass)
        if shortcut is not None:
            ops.add_bf16(dx, shortcut)
        if strided is not None:
            ops.add_strided2(dx, strided)
        return dx


class ResNetRT:
    """Whole-network runtime for ResNet (resnet.py:158-245) and ResNetCifar
    (resnetforcifar.py:27-108): stem [+ maxpool] + 4 stages + avgpool + fc."""

    def __init__(self, model, has_maxpool):
        self.model = model
        self.has_maxpool = has_maxpool
        self.stem = ConvBN(model.conv1, ACT_RELU)
        self.blocks = []
        for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
            for blk in layer:
                self.blocks.append(ResidualBlockRT(blk))
        self.fc = model.fc
        self.fc_w_bf16 = None
        self.fc_b_pad = None
        self.fc_version = None
        self.sink = GradSink()
        self.tape = None

    def units(self):
        us = [self.stem]
        for b in self.blocks:
            us += b.all_units()
        return us

    def prep(self):
        for u in self.units():
            u.prep()
        w = self.fc.weight
        ver = (w.data_ptr(), w._version)
        if self.fc_w_bf16 is None or ver != self.fc_version:
            npad = _round_up(w.shape[0], 8)
            if self.fc_w_bf16 is None:
                self.fc_w_bf16 = torch.zeros(npad, w.shape[1], device=w.device, dtype=torch.bfloat16)
                self.fc_b_pad = torch.zeros(npad, device=w.device)
            ops.cast_bf16(w.detach(), self.fc_w_bf16[:w.shape[0]])
            self.fc_version = ver
        self.fc_b_pad[:w.shape[0]].copy_(self.fc.bias.detach())

    # The network is run as three stages (stem / residual blocks / head) so that tests can drive
    # each stage with the oracle's tensors (tests/test_resnet_gpu.py, teacher-forced parity).
    def stem_forward(self, x, tape, training):
        a = self.stem.forward(x, tape['stem'], training)
        if self.has_maxpool:
            tape['pool_in_hw'] = (a.shape[1], a.shape[2])
            a, tape['argmax'] = ops.maxpool3x3s2_fwd(a)
        return a

    def head_forward(self, a, tape):
        tape['feat_hw'] = (a.shape[1], a.shape[2])
        pooled = ops.avgpool_fwd(a)
        tape['pooled'] = pooled
        ncls = self.fc.weight.shape[0]
        logits = ops.linear_fwd(pooled, self.fc_w_bf16, bias=self.fc_b_pad, out_f32=True)
        if logits.shape[1] != ncls:
            logits = logits[:, :ncls].contiguous()
        return logits

    def forward(self, x, training, keep_tape):
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
        x = x.contiguous()
        self.prep()
        tape = {'stem': {}, 'blocks': [dict() for _ in self.blocks]}
        a = self.stem_forward(x, tape, training)
        for b, t in zip(self.blocks, tape['blocks']):
            a = b.forward(a, t, training)
        logits = self.head_forward(a, tape)
        self.tape = tape if keep_tape else None
        return logits

    def head_backward(self, dlogits, tape):
        """dlogits fp32 [B, num_classes] -> gradient w.r.t. the last feature map (NHWC bf16)."""
        sink = self.sink
        ncls, feat = self.fc.weight.shape
        npad = self.fc_w_bf16.shape[0]
        dlogits = dlogits.contiguous().float()
        bbuf, bacc = sink.begin(self.fc.bias)
        ops.colsum(dlogits, bbuf, accumulate=bacc)
        sink.done(self.fc.bias, bbuf)
        dl = torch.zeros(dlogits.shape[0], npad, device=dlogits.device, dtype=torch.bfloat16)
        if npad == ncls:
            ops.cast_bf16(dlogits, dl)
        else:
            dl[:, :ncls] = dlogits.to(torch.bfloat16)
        wbuf, wacc = sink.begin(self.fc.weight)
        part = ops.linear_wgrad(dl, tape['pooled'])
        if npad == ncls:
            ops.reduce_partials(part, wbuf, accumulate=wacc)
        else:
            tmp = torch.empty(npad, feat, device=dl.device)
            ops.reduce_partials(part, tmp)
            wbuf.copy_(tmp[:ncls] + (wbuf if wacc else 0))
        sink.done(self.fc.weight, wbuf)
        dpooled = ops.linear_dgrad(dl, self.fc_w_bf16)
        h, w = tape['feat_hw']
        return ops.avgpool_bwd(dpooled, h, w)

    def stem_backward(self, da, tape):
        if self.has_maxpool:
            ph, pw = tape['pool_in_hw']
            da = ops.maxpool3x3s2_bwd(da, tape['argmax'], ph, pw)
        dy, _ = self.stem.bn_bwd(da, tape['stem'], self.sink)
        self.stem.conv_bwd(dy, tape['stem'], self.sink, need_dx=False)

    def backward(self, dlogits):
        tape, sink = self.tape, self.sink
        assert tape is not None, 'backward called without a training forward'
        self.tape = None
        da = self.head_backward(dlogits, tape)
        for b, t in zip(reversed(self.blocks), reversed(tape['blocks'])):
            da = b.backward(da, t, sink)
        self.stem_backward(da, tape)
        if sink.on_backward_end is not None:
            sink.on_backward_end()


class _NetFunction(torch.autograd.Function):
    """Couples the runtime to autograd: the criterion (torch) differentiates the logits, this
    node receives dlogits and runs the whole backward pass on our kernels.  Parameter gradients
    are produced as a side effect (GradSink), so no parameter is an autograd input."""

    @staticmethod
    def forward(ctx, x, anchor, rt):
        ctx.rt = rt
        return rt.forward(x, True, True)

    @staticmethod
    def backward(ctx, dlogits):
        ctx.rt.backward(dlogits)
        return None, None, None


def run_network(rt, x, training):
    if training and torch.is_grad_enabled():
        anchor = torch.zeros((), device=x.device, requires_grad=True)
        return _NetFunction.apply(x, anchor, rt)
    with torch.no_grad():
        return rt.forward(x, training, False)
