"""Execution engines for G_NET and D_NET64/128/256: hand-scheduled forward and backward over the sg2b200 kernels.

The nn.Module classes in model.py only hold parameters/buffers (reference names and order); these engines read
them, run NHWC-bf16 pipelines on the CUDA kernels and hand gradients back to autograd. Reference behaviour
mirrored: StackGAN_v2/model.py:112-354 (G side), 358-551 (D side).
"""
import torch

from . import ops
from .ops import ACT_GLU, ACT_LRELU, ACT_NONE, CONV3, CONV4S2, GEMM, STEM, UPCONV


def _pad_to(n, m):
    return -(-n // m) * m


class ConvOperand:
    """bf16 operand packs (fprop + dgrad) and the fp32 wgrad accumulator of one conv weight."""

    def __init__(self, kind, weight, cout_pad=None, cin_pad=None, precise=False):
        self.kind = kind
        self.weight = weight
        self.precise = precise      # fp32-accurate mode: packs are 3-plane bf16 splits of the fp32 packs (ops.split3)
        self.Cout = weight.shape[0]
        self.Cin = 48 if kind == STEM else weight.shape[1]
        self.CoP = cout_pad or self.Cout
        self.CiP = cin_pad or self.Cin
        self._key = None
        self.wpk = self.wpkT = self.dwpk = None
        # Module-API mode: the reference writes weights through `p.data` (load_params' EMA swap for snapshot images and
        # save_model, trainer.py:78-80,256,592-601; weights_init), which bumps no version counter. Those forwards all run
        # under torch.no_grad(), so with auto_refresh a no_grad forward always re-packs and leaves the cache invalid for
        # the next (training) forward. The fused trainer owns its weights and switches this off.
        self.auto_refresh = True
        # algorithmic / executed FLOP ratio of the padded operand (image heads pad 3 -> 32, stems pad 48 -> 64)
        self.flop_scale = (self.Cout * self.Cin) / float(self.CoP * self.CiP)

    # ---- OHWI fast path (FusedTrainer's FlatBucket): the fp32 master of a conv weight is stored [Cout][kh][kw][Cin];
    # `weight._sg2_ohwi` is that contiguous master, `_sg2_wpk` its bf16 mirror (kept current by the Adam kernel) and
    # `_sg2_dw` the matching slice of the flat gradient bucket. For CONV3x3 / CONV4x4S2 without channel padding the
    # mirror IS the fprop operand and the packed wgrad accumulator IS the gradient: no pack, no unpack.
    def _ohwi(self):
        return getattr(self.weight, "_sg2_ohwi", None)

    def _direct(self):
        return (not self.precise and self._ohwi() is not None and self.kind in (CONV3, CONV4S2) and self.CoP == self.Cout
                and self.CiP == self.Cin and self.Cout % 8 == 0 and self.Cin % 8 == 0)

    def _pack_key(self):
        """(cache key, force): force = re-pack regardless of the key and do not cache (see auto_refresh)."""
        w = self.weight
        # _sg2_version: bumped by FlatBucket.adam(), whose kernel updates the weights behind torch's back, and by
        # utils.load_params / invalidate_packs for writes through `.data`
        key = (w.data_ptr(), w._version, getattr(w, "_sg2_version", 0), w.device)
        return key, (self.auto_refresh and not torch.is_grad_enabled())

    def packs(self):
        w = self.weight
        key, force = self._pack_key()
        oh = self._ohwi()
        if self._direct():
            if force or key != self._key:
                if force or self._key is None or key[1] != self._key[1] or getattr(w, "_sg2_mirror_stale", False):
                    ops.f32_to_bf16(oh, out=w._sg2_wpk)      # torch-side write to the master: refresh the bf16 mirror
                    w._sg2_mirror_stale = False
                self.wpk = w._sg2_wpk
                if self.wpkT is None or self.wpkT.device != w.device:
                    _, s2 = ops.pack_shapes(self.kind, self.CoP, self.CiP)
                    self.wpkT = torch.empty(s2, device=w.device, dtype=torch.bfloat16)
                ops.pack_transpose(self.kind, self.wpk, self.wpkT, self.Cout, self.Cin)
                self._key = None if force else key
            return self.wpk, self.wpkT
        if self.precise:
            if force or key != self._key:
                self.wpk, self.wpkT = ops.pack_weights_split3(self.kind, w.detach() if oh is None else oh, self.Cout,
                                                              self.Cin, self.CoP, self.CiP, ohwi=oh is not None)
                self._key = None if force else key
            return self.wpk, self.wpkT
        if force or key != self._key:
            s1, s2 = ops.pack_shapes(GEMM if self.kind == STEM else self.kind, self.CoP, self.CiP)
            if self.wpk is None or self.wpk.device != w.device:
                self.wpk = torch.empty(s1, device=w.device, dtype=torch.bfloat16)
                self.wpkT = torch.empty(s2, device=w.device, dtype=torch.bfloat16)
            ops.pack_weights(self.kind, w.detach() if oh is None else oh, self.wpk, self.wpkT, self.Cout, self.Cin,
                             self.CoP, self.CiP, ohwi=oh is not None)
            self._key = None if force else key
        return self.wpk, self.wpkT

    def wgrad_begin(self, device, prezeroed=False):
        """prezeroed: the caller cleared the whole flat gradient bucket (direct accumulation needs no fill here)."""
        self._first = True      # reproducible mode: the first pass writes the accumulator, no clearing needed
        if self._direct():
            self.dwpk = self.weight._sg2_dw
            if not prezeroed and not ops.DETERMINISTIC:
                self.dwpk.zero_()
            return
        k = GEMM if self.kind == STEM else self.kind
        shape = (self.CoP, ops.JOBS[k], self.CiP)
        if self.dwpk is None or self.dwpk.device != device or self.dwpk.shape != shape:
            self.dwpk = torch.empty(shape, device=device, dtype=torch.float32)
        if not ops.DETERMINISTIC:
            self.dwpk.zero_()

    def wgrad_add(self, x, dy):
        """dwpk (=|+=) dy^T im2col(x) (tcgen05 wgrad kernel; ordered slab reduction, or fp32 red.global.add)."""
        ops.conv_wgrad(GEMM if self.kind == STEM else self.kind, x, dy, self.dwpk, flop_scale=self.flop_scale,
                       first=self._first)
        self._first = False

    def wgrad_finish(self, out=None):
        """packed fp32 accumulator -> fp32 gradient in the master's layout (written into `out` if given)."""
        oh = self._ohwi()
        if oh is not None:
            w = self.weight
            if not self._direct():                             # direct: accumulated in place in the gradient bucket
                ops.unpack_wgrad(self.kind, self.dwpk, w._sg2_dw, self.Cout, self.Cin, self.CoP, self.CiP, False, ohwi=True)
            if out is not None:
                return out                                     # = the OIHW-shaped view of that bucket slice
            Co, Ci, kh, kw = w.shape        # module API (autograd owns the result): hand out a private OIHW copy
            return w._sg2_dw.view(Co, kh, kw, Ci).permute(0, 3, 1, 2).contiguous()
        g = torch.empty_like(self.weight) if out is None else out
        ops.unpack_wgrad(self.kind, self.dwpk, g, self.Cout, self.Cin, self.CoP, self.CiP, False)
        return g


class GradSink:
    """Collects parameter gradients of one or several backward passes (e.g. D's real/wrong/fake passes).

    Conv weight gradients accumulate in the packed fp32 buffers and are unpacked once in finish(); the small
    BN / linear / logit gradients accumulate in place. `views` optionally maps parameter -> preallocated
    fp32 tensor (a slice of a flat gradient bucket) that receives the result."""

    def __init__(self, views=None, side_stream=None, prezeroed=False, on_ready=None, opt_stream=None):
        self.g = {}
        self.views = views or {}
        self.pending = []
        self.prezeroed = prezeroed     # the flat gradient bucket behind `views` was cleared by the caller
        # on_ready(weight): called (on the stream the wgrad ran on) as soon as a conv weight's gradient is complete, i.e.
        # after its single wgrad launch of this sink — lets the caller all-reduce and Adam-update that layer while the
        # rest of the backward pass is still running. Only valid when every conv sees exactly one backward pass.
        self.on_ready = on_ready
        # on_repack(op): optional; called on the side stream once the layer's weight has been updated (on_ready) AND its
        # own dgrad — the last reader of the old dgrad operand — has been issued: the caller re-packs the layer's bf16
        # operands there, off the critical path, instead of lazily at the next forward.
        self.on_repack = None
        # Weight gradients hang off the backward chain (only the optimiser consumes them), so they can run on a
        # side stream next to the dgrad / BatchNorm-backward chain; finish() joins.
        self.side = side_stream
        # optional second side stream for what follows a layer's wgrad (unpack, all-reduce, Adam, operand re-pack): the
        # next layer's wgrad then runs beside the previous layer's optimiser work instead of behind it
        self.opt = opt_stream if side_stream is not None else None
        self.keep = []
        self.started = set()   # conv operands whose accumulator has been opened (several backward passes may add to it)

    def slot(self, p):
        """(tensor, accumulate_flag) for a small gradient written by a kernel."""
        if p in self.g:
            return self.g[p], True
        t = self.views.get(p)
        if t is None:
            t = torch.empty_like(p, dtype=torch.float32)
        self.g[p] = t
        return t, False

    def zero_slot(self, p):
        """Tensor for `+=` kernels: zeroed on first use."""
        t, acc = self.slot(p)
        if not acc:
            t.zero_()
        return t

    def conv(self, op, x, dy):
        if self.side is None:
            self._wgrad(op, x, dy)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.keep.append((x, dy))              # keep the operands alive until the side stream has consumed them
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self._wgrad(op, x, dy)

    def dgrad_done(self, op):
        """The layer's dgrad (if any) has been issued on the current stream."""
        if self.on_repack is None or self.on_ready is None:
            return
        if self.side is None:
            self.on_repack(op)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        st = self.opt if self.opt is not None else self.side
        with torch.cuda.stream(st):
            st.wait_event(ev)
            self.on_repack(op)

    def _wgrad(self, op, x, dy):
        if self.on_ready is not None:
            # the layer's LAST backward pass of this sink (earlier passes, run with on_ready = None, only accumulated)
            if op not in self.started:
                op.wgrad_begin(x.device, self.prezeroed)
                self.started.add(op)
            if op in self.pending:
                self.pending.remove(op)
            op.wgrad_add(x, dy)
            if self.opt is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())       # = the wgrad side stream
                with torch.cuda.stream(self.opt):
                    self.opt.wait_event(ev)
                    self.g[op.weight] = op.wgrad_finish(self.views.get(op.weight))
                    self.on_ready(op.weight)
                return
            self.g[op.weight] = op.wgrad_finish(self.views.get(op.weight))
            self.on_ready(op.weight)
            return
        if op not in self.started:
            op.wgrad_begin(x.device, self.prezeroed)
            self.started.add(op)
        if op not in self.pending:
            self.pending.append(op)
        op.wgrad_add(x, dy)

    def finish(self):
        if self.side is None:
            for op in self.pending:
                self.g[op.weight] = op.wgrad_finish(self.views.get(op.weight))
        else:
            with torch.cuda.stream(self.side):
                for op in self.pending:
                    self.g[op.weight] = op.wgrad_finish(self.views.get(op.weight))
                ev = torch.cuda.Event()
                ev.record(self.side)
            torch.cuda.current_stream().wait_event(ev)
            if self.opt is not None:
                ev2 = torch.cuda.Event()
                ev2.record(self.opt)
                torch.cuda.current_stream().wait_event(ev2)
        self.pending = []
        self.keep = []
        return self.g


class ConvBlock:
    """conv (3x3 | fused-upsample 3x3 | 4x4 s2) -> [BatchNorm] -> GLU | LeakyReLU | (+residual)."""

    def __init__(self, kind, conv, bn, act, precise=False):
        self.kind, self.conv, self.bn, self.act = kind, conv, bn, act
        self.op = ConvOperand(kind, conv.weight, precise=precise)

    def fwd(self, x, training, residual=None, groups=1):
        """groups > 1: the batch is `groups` equal sub-batches with separate BatchNorm statistics (the batched
        real / wrong / fake passes of train_Dnet, trainer.py:390-392)."""
        wpk, _ = self.op.packs()
        bn = self.bn
        if bn is None:
            y = ops.conv_fprop(self.kind, x, wpk, self.op.Cout)
            return ops.bn_act_fwd(y, None, None, self.act, residual), (x, y, None, None)
        gamma, beta = bn.weight.detach(), bn.bias.detach()
        if training:
            # batch statistics come out of the conv epilogue (or the split-K fp32->bf16 pass): no pass over y
            st = ops.bn_stats32(self.op.Cout, x.device, groups)
            y, _ = ops.conv_fprop(self.kind, x, wpk, self.op.Cout, stats=st, groups=groups)
            out, mean, rstd = ops.bn_act_fwd(y, gamma, beta, self.act, residual, stats=st, groups=groups,
                                             running=(bn.running_mean, bn.running_var, bn.num_batches_tracked))
        else:
            y = ops.conv_fprop(self.kind, x, wpk, self.op.Cout)
            mean, rstd = ops.bn_eval_stats(bn.running_mean, bn.running_var)
            out = ops.bn_act_fwd(y, gamma, beta, self.act, residual, mean=mean, rstd=rstd)
        return out, (x, y, mean, rstd)

    def bwd(self, saved, dout, sink, need_dx=True, need_w=True, groups=1, epi=None):
        """epi = (src, ops.EPI_ADD | ops.EPI_LRELU_MASK): folded into the dgrad epilogue (see ops.conv_dgrad)."""
        x, y, mean, rstd = saved
        if self.bn is not None:
            if need_w:
                (dg, acc), (db, _) = sink.slot(self.bn.weight), sink.slot(self.bn.bias)
            else:
                dg = db = None
                acc = False
            dy = ops.bn_act_bwd(y, dout, mean, rstd, self.bn.weight.detach(), self.bn.bias.detach(), self.act,
                                dg, db, acc, groups=groups)
        elif self.act == ACT_LRELU:
            dy = ops.lrelu_bwd(y, dout)
        else:
            dy = dout
        if need_w:
            sink.conv(self.op, x, dy)
        dx = None
        if need_dx:
            _, wpkT = self.op.packs()
            B, H, W, Cin = x.shape
            dx = ops.conv_dgrad(self.kind, dy, wpkT, B, H, W, Cin, epi=epi)
        if need_w:
            sink.dgrad_done(self.op)
        return dx


class JointOperand:
    """Operands of a generator jointConv (conv3x3 on cat(c_code broadcast, h), model.py:274-279) with the c_code
    channels folded away: the tensor-core kernels only see the h part of the weight, [Cout][9][Ch]; the first E input
    channels act through a per-sample bias (forward) and two small reductions (backward). Same interface as
    ConvOperand, so GradSink / the per-layer optimiser treat it like any other conv weight."""
    kind = CONV3

    def __init__(self, weight, E):
        self.weight = weight
        self.E = E
        self.Cout, self.CinFull = weight.shape[0], weight.shape[1]
        self.Cin = self.CinFull - E
        self.CoP, self.CiP = self.Cout, self.Cin
        self.flop_scale = 1.0
        self._key = None
        self._full = ConvOperand(CONV3, weight)      # bf16 [Cout][9][CinFull] of the whole weight (mirror or pack)
        self.wpk = self.wpkT = self.dwpk = None
        self.cparts = []
        self.auto_refresh = True                      # see ConvOperand

    _pack_key = ConvOperand._pack_key

    def w_f32(self):
        """(fp32 master, so, se, st): element (o, e, tap) of the full weight at flat[o*so + e*se + tap*st]."""
        oh = getattr(self.weight, "_sg2_ohwi", None)
        if oh is not None:
            return oh, 9 * self.CinFull, 1, self.CinFull
        return self.weight.detach(), 9 * self.CinFull, 9, 1

    def packs(self):
        w = self.weight
        key, force = self._pack_key()
        if force or key != self._key:
            oh = getattr(w, "_sg2_ohwi", None)
            if oh is not None:
                if force or self._key is None or key[1] != self._key[1] or getattr(w, "_sg2_mirror_stale", False):
                    ops.f32_to_bf16(oh, out=w._sg2_wpk)      # torch-side write to the master: refresh the bf16 mirror
                    w._sg2_mirror_stale = False
                full = w._sg2_wpk
            else:
                self._full.auto_refresh = self.auto_refresh
                full, _ = self._full.packs()
            if self.wpk is None or self.wpk.device != w.device:
                self.wpk = torch.empty((self.Cout, 9, self.Cin), device=w.device, dtype=torch.bfloat16)
                self.wpkT = torch.empty((self.Cin, 9, self.Cout), device=w.device, dtype=torch.bfloat16)
            self.wpk.copy_(full.view(self.Cout, 9, self.CinFull)[:, :, self.E:])
            ops.pack_transpose(CONV3, self.wpk, self.wpkT, self.Cout, self.Cin)
            self._key = None if force else key
        return self.wpk, self.wpkT

    def wgrad_begin(self, device, prezeroed=False):
        shape = (self.Cout, 9, self.Cin)
        if self.dwpk is None or self.dwpk.device != device:
            self.dwpk = torch.empty(shape, device=device, dtype=torch.float32)
        if not ops.DETERMINISTIC:
            self.dwpk.zero_()
        self._first = True

    def wgrad_add(self, x, dy):
        ops.conv_wgrad(CONV3, x, dy, self.dwpk, first=self._first)
        self._first = False

    def wgrad_finish(self, out=None):
        """h part: the packed accumulator goes into the [.., E:] channels of the gradient; c part: c^T S."""
        w = self.weight
        oh = getattr(w, "_sg2_ohwi", None)
        if oh is not None:
            g = w._sg2_dw.view(self.Cout, 9, self.CinFull)
            g[:, :, self.E:].copy_(self.dwpk)
            dst, strides = w._sg2_dw, (9 * self.CinFull, 1, self.CinFull)
        else:
            g = torch.empty_like(w, dtype=torch.float32) if out is None else out
            g[:, self.E:].copy_(self.dwpk.view(self.Cout, 3, 3, self.Cin).permute(0, 3, 1, 2))
            dst, strides = g, (9 * self.CinFull, 9, 1)
        for k, (c, S) in enumerate(self.cparts):
            ops.joint_c_bwd(S, c, (dst,) + strides, dw=dst, dw_accumulate=k > 0)
        self.cparts = []
        if oh is None:
            return g
        if out is not None:
            return out
        Co, Ci, kh, kw = w.shape
        return w._sg2_dw.view(Co, kh, kw, Ci).permute(0, 3, 1, 2).contiguous()


class JointBlock:
    """jointConv of NEXT_STAGE_G: conv3x3(cat(c_code, h)) -> BatchNorm -> GLU without materialising the concatenation."""

    def __init__(self, conv, bn, E, precise=False):
        self.conv, self.bn, self.act = conv, bn, ACT_GLU
        self.E = E
        # fp32-accurate mode: the concatenation is materialised and convolved with the whole weight, exactly like the
        # reference (model.py:274-279); the folding below is an optimisation of the bf16 path
        self.full = ConvBlock(CONV3, conv, bn, ACT_GLU, precise=True) if precise else None
        self.op = self.full.op if precise else JointOperand(conv.weight, E)

    def fwd(self, c, h, training):
        if self.full is not None:
            out, sv = self.full.fwd(ops.concat_c(c, h), training)
            return out, sv + (c,)
        op, bn = self.op, self.bn
        wpk, _ = op.packs()
        bias9 = ops.joint_bias(c, op.w_f32(), op.Cout)
        gamma, beta = bn.weight.detach(), bn.bias.detach()
        if training:
            st = ops.bn_stats32(op.Cout, h.device)
            y, _ = ops.conv_fprop(CONV3, h, wpk, op.Cout, stats=st, bias9=bias9)
            out, mean, rstd = ops.bn_act_fwd(y, gamma, beta, self.act, stats=st,
                                             running=(bn.running_mean, bn.running_var, bn.num_batches_tracked))
        else:
            y = ops.conv_fprop(CONV3, h, wpk, op.Cout, bias9=bias9)
            mean, rstd = ops.bn_eval_stats(bn.running_mean, bn.running_var)
            out = ops.bn_act_fwd(y, gamma, beta, self.act, mean=mean, rstd=rstd)
        return out, (h, y, mean, rstd, c)

    def bwd(self, saved, dout, sink, dc):
        """-> dh; dc (B, E) fp32 += the c_code gradient of this layer."""
        if self.full is not None:
            dcat = self.full.bwd(saved[:4], dout, sink)
            return ops.concat_c_bwd(dcat, self.E, dc)
        h, y, mean, rstd, c = saved
        op = self.op
        (dg, acc), (db, _) = sink.slot(self.bn.weight), sink.slot(self.bn.bias)
        dy = ops.bn_act_bwd(y, dout, mean, rstd, self.bn.weight.detach(), self.bn.bias.detach(), self.act, dg, db, acc)
        S = ops.joint_tap_sums(dy)
        ops.joint_c_bwd(S, c, op.w_f32(), dc=dc)
        op.cparts.append((c, S))
        # S (and c) are read again by wgrad_finish() on the wgrad side stream after this function has returned: keep
        # them allocated until GradSink.finish() so the main stream's allocator cannot hand the memory out in between
        sink.keep.append((c, S))
        sink.conv(op, h, dy)
        _, wpkT = op.packs()
        B, H, W, Ch = h.shape
        dh = ops.conv_dgrad(CONV3, dy, wpkT, B, H, W, Ch)
        sink.dgrad_done(op)
        return dh


class HeadBlock:
    """GET_IMAGE_G (model.py:287-298): conv3x3 C->3 + tanh; output NCHW fp32 image."""
    CP = 16

    def __init__(self, conv, precise=False):
        self.conv = conv
        self.precise = precise
        self.op = ConvOperand(CONV3, conv.weight, cout_pad=self.CP, precise=precise)

    def fwd(self, h):
        wpk, _ = self.op.packs()
        y = ops.conv_fprop(CONV3, h, wpk, self.CP, flop_scale=self.op.flop_scale)
        B, H, W, _ = h.shape
        return ops.head_tanh_fwd(y, B, H, W), h

    def bwd(self, h, img, dimg, sink):
        dy = ops.head_tanh_bwd(dimg, img, self.CP, f32=self.precise)
        sink.conv(self.op, h, dy)
        _, wpkT = self.op.packs()
        B, H, W, C = h.shape
        dx = ops.conv_dgrad(CONV3, dy, wpkT, B, H, W, C, flop_scale=self.op.flop_scale)
        sink.dgrad_done(self.op)
        return dx


def _acc(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return ops.add_bf16(a, b)


# =============================================================================================== generator
class GEngine:
    def __init__(self, net, cfg, precise=False):
        self.net = net
        self.precise = pr = precise     # fp32-accurate mode (fp32 activations, 3-way bf16 split convolutions)
        self.E = cfg.GAN.EMBEDDING_DIM
        self.Z = cfg.GAN.Z_DIM
        self.branches = cfg.TREE.BRANCH_NUM
        self.ngf = cfg.GAN.GF_DIM * 16
        h1 = net.h_net1
        self.ups1 = [ConvBlock(UPCONV, getattr(h1, f"upsample{i}")[1], getattr(h1, f"upsample{i}")[2], ACT_GLU, pr)
                     for i in (1, 2, 3, 4)]
        self.heads = [HeadBlock(net.img_net1.img[0], pr)]
        self.stages = []
        for s in range(2, self.branches + 1):
            hn = getattr(net, f"h_net{s}")
            joint = JointBlock(hn.jointConv[0], hn.jointConv[1], self.E, pr)
            res = [(ConvBlock(CONV3, r.block[0], r.block[1], ACT_GLU, pr), ConvBlock(CONV3, r.block[3], r.block[4], ACT_NONE, pr))
                   for r in hn.residual]
            up = ConvBlock(UPCONV, hn.upsample[1], hn.upsample[2], ACT_GLU, pr)
            self.stages.append((joint, res, up))
            self.heads.append(HeadBlock(getattr(net, f"img_net{s}").img[0], pr))

    def params(self):
        return [p for p in self.net.parameters()]

    def set_auto_refresh(self, flag):
        for op in self.conv_ops():
            op.auto_refresh = flag

    def conv_ops(self):
        ops_ = [b.op for b in self.ups1] + [h.op for h in self.heads]
        for joint, res, up in self.stages:
            ops_ += [joint.op, up.op] + [b.op for pair in res for b in pair]
        return ops_

    # ------------------------------------------------------------------ forward
    def forward(self, z, emb, eps, training, on_mu=None, on_img=None):
        """on_mu(mu): called as soon as CA_NET has produced mu (the discriminators' conditioning, trainer.py:383), so that
        work which needs mu but not the fake images can be issued while the rest of the generator runs.
        on_img(i, img): called as soon as stage i's image has been issued (the smaller discriminators' updates only need
        their own scale: they can start while the later generator stages are still running)."""
        net = self.net
        B = z.shape[0]
        T = {}
        ca = net.ca_net.fc
        T["emb"], T["z"], T["eps"] = emb, z, eps
        T["fc_ca"] = ops.linear_fwd(emb, None, ca.weight.detach(), ca.bias.detach(), False)
        mu, logvar, c = ops.ca_glu_reparam_fwd(T["fc_ca"], eps)
        T["c"] = c
        if on_mu is not None:
            on_mu(mu)
        fc, bn = net.h_net1.fc[0], net.h_net1.fc[1]
        h32 = ops.linear_fwd(c, z, fc.weight.detach(), None, False)                    # (B, ngf*32) fp32
        if training:
            st = ops.bn_stats32(h32.shape[1], h32.device)
            h = ops.f32_to_bf16_stats(h32, st, keep_f32=self.precise)
            g, mean, rstd = ops.bn_act_fwd(h, bn.weight.detach(), bn.bias.detach(), ACT_GLU, stats=st,
                                           running=(bn.running_mean, bn.running_var, bn.num_batches_tracked))
        else:
            h = h32 if self.precise else ops.f32_to_bf16(h32)
            mean, rstd = ops.bn_eval_stats(bn.running_mean, bn.running_var)
            g = ops.bn_act_fwd(h, bn.weight.detach(), bn.bias.detach(), ACT_GLU, mean=mean, rstd=rstd)  # CHW order
        T["fc"] = (h, mean, rstd)
        x = ops.chw_hwc(g, B, self.ngf, 16, True).view(B, 4, 4, self.ngf)
        T["ups1"] = []
        for blk in self.ups1:
            x, sv = blk.fwd(x, training)
            T["ups1"].append(sv)
        imgs = []
        img, hsv = self.heads[0].fwd(x)
        imgs.append(img)
        if on_img is not None:
            on_img(0, img)
        T["heads"] = [hsv]
        T["stages"] = []
        for si, (joint, res, up) in enumerate(self.stages):
            x, sj = joint.fwd(c, x, training)
            sres = []
            for (b0, b1) in res:
                mid, s0 = b0.fwd(x, training)
                x, s1 = b1.fwd(mid, training, residual=x)
                sres.append((s0, s1))
            x, su = up.fwd(x, training)
            T["stages"].append((sj, sres, su))
            img, hsv = self.heads[si + 1].fwd(x)
            imgs.append(img)
            if on_img is not None:
                on_img(si + 1, img)
            T["heads"].append(hsv)
        T["imgs"] = imgs
        return imgs, mu, logvar, T

    # ------------------------------------------------------------------ backward
    def backward(self, T, dimgs, dmu, dlogvar, sink=None):
        """-> dict {parameter: fp32 grad}. dimgs[i] may be None."""
        net = self.net
        own = sink is None
        grads = sink = GradSink() if own else sink
        c = T["c"]
        B = c.shape[0]
        dc = torch.zeros_like(c)
        dx = None   # gradient flowing into the h_code of the current stage (NHWC bf16)
        for si in range(len(self.stages), -1, -1):
            head = self.heads[si]
            if dimgs[si] is not None:
                dx = _acc(dx, head.bwd(T["heads"][si], T["imgs"][si], dimgs[si].contiguous(), grads))
            else:
                sink.zero_slot(head.conv.weight)
            if si == 0:
                break
            joint, res, up = self.stages[si - 1]
            sj, sres, su = T["stages"][si - 1]
            if dx is None:
                raise RuntimeError("sg2b200: no gradient reached generator stage %d" % (si + 1))
            dx = up.bwd(su, dx, grads)
            for (b0, b1), (s0, s1) in zip(reversed(res), reversed(sres)):
                dmid = b1.bwd(s1, dx, grads)
                dx = b0.bwd(s0, dmid, grads, epi=(dx, ops.EPI_ADD))     # + the skip branch's gradient
            dx = joint.bwd(sj, dx, grads, dc)
        for blk, sv in zip(reversed(self.ups1), reversed(T["ups1"])):
            dx = blk.bwd(sv, dx, grads)
        # INIT_STAGE_G fc: NHWC -> CHW feature order -> BN1d+GLU backward -> linear
        fc, bn = net.h_net1.fc[0], net.h_net1.fc[1]
        h, mean, rstd = T["fc"]
        dg = ops.chw_hwc(dx.reshape(B, -1), B, self.ngf, 16, False)
        (dgam, acc), (dbet, _) = sink.slot(bn.weight), sink.slot(bn.bias)
        dh = ops.bn_act_bwd(h, dg, mean, rstd, bn.weight.detach(), bn.bias.detach(), ACT_GLU, dgam, dbet, acc)
        dw, acc = sink.slot(fc.weight)
        ops.linear_bwd_w(dh, c, T["z"], dw, None, acc)
        dc.add_(ops.linear_bwd_x(dh, fc.weight.detach(), self.E))     # (B, E) fp32: plumbing-sized
        ca = net.ca_net.fc
        dfc = ops.ca_glu_reparam_bwd(T["fc_ca"], T["eps"], dmu, dlogvar, dc)
        (dw, acc), (db, _) = sink.slot(ca.weight), sink.slot(ca.bias)
        ops.linear_bwd_w(dfc, T["emb"], None, dw, db, acc)
        return sink.finish() if own else None


# =============================================================================================== discriminators
class StemBlock:
    """encode_image_by_16times[0:2] (model.py:383-384): conv4x4 s2 3->ndf (no BN) + LeakyReLU, as im2col + GEMM."""

    def __init__(self, conv, precise=False):
        self.conv = conv
        self.precise = precise
        self.op = ConvOperand(STEM, conv.weight, cin_pad=64, precise=precise)

    def fwd(self, img, col=None):
        """col: optional precomputed im2col rows of `img` (the G step reuses the fake third of the D update's); `img`
        may then be just the (B, S) pair."""
        B, S = img if isinstance(img, tuple) else (img.shape[0], img.shape[2])
        if col is None:
            col = ops.stem_im2col(img, f32=self.precise)
        wpk, _ = self.op.packs()
        # LeakyReLU runs in the GEMM epilogue; backward only needs the sign, and sign(lrelu(y)) == sign(y).
        # The im2col rows are presented as a (B, S/2, S/2, 64) NHWC tensor: a 1x1 conv on the tile-resident kernel.
        out = ops.conv_fprop(GEMM, col.view(B, S // 2, S // 2, 64), wpk, self.op.Cout, flop_scale=self.op.flop_scale,
                             act=ACT_LRELU)
        return out, (col, out, B, S)

    def bwd(self, saved, dout, sink, need_dimg, need_w=True, masked=False):
        """masked: dout already carries the LeakyReLU backward (folded into the producing dgrad's epilogue)."""
        col, y, B, S = saved
        dy = (dout if masked else ops.lrelu_bwd(y, dout)).view(1, 1, -1, self.op.Cout)
        if need_w:
            sink.conv(self.op, col, dy)
        dimg = None
        if need_dimg:
            _, wpkT = self.op.packs()
            dcol = ops.conv_dgrad(GEMM, dy, wpkT, 1, 1, col.shape[2], 64, flop_scale=self.op.flop_scale)
            dimg = ops.stem_col2im(dcol, B, S)
        if need_w:
            sink.dgrad_done(self.op)
        return dimg


class DEngine:
    def __init__(self, net, cfg, precise=False):
        self.net = net
        self.precise = pr = precise
        self.E = cfg.GAN.EMBEDDING_DIM
        s16 = net.img_code_s16
        self.stem = StemBlock(s16[0], pr)
        self.trunk = [ConvBlock(CONV4S2, s16[2], s16[3], ACT_LRELU, pr), ConvBlock(CONV4S2, s16[5], s16[6], ACT_LRELU, pr),
                      ConvBlock(CONV4S2, s16[8], s16[9], ACT_LRELU, pr)]
        for name, kind in (("img_code_s32", CONV4S2), ("img_code_s64", CONV4S2), ("img_code_s32_1", CONV3),
                           ("img_code_s64_1", CONV3), ("img_code_s64_2", CONV3)):
            if hasattr(net, name):
                m = getattr(net, name)
                self.trunk.append(ConvBlock(kind, m[0], m[1], ACT_LRELU, pr))
        self.joint = ConvBlock(CONV3, net.jointConv[0], net.jointConv[1], ACT_LRELU, pr)

    def params(self):
        return [p for p in self.net.parameters()]

    def set_auto_refresh(self, flag):
        for op in self.conv_ops():
            op.auto_refresh = flag

    def conv_ops(self):
        return [self.stem.op, self.joint.op] + [b.op for b in self.trunk]

    def forward(self, img, c, training, out_cond=None, out_uncond=None, groups=1, stem_col=None, want_features=True):
        """groups > 1: img / c hold `groups` equal sub-batches that the reference runs as separate D passes (separate
        BatchNorm batches, trainer.py:390-392); everything else is per sample, so one pass over the concatenation
        gives the same result."""
        T = {"groups": groups}
        x, T["stem"] = self.stem.fwd(img, stem_col)
        T["trunk"] = []
        for blk in self.trunk:
            x, sv = blk.fwd(x, training, groups=groups)
            T["trunk"].append(sv)
        T["x_code"] = x
        # x_immediate (model.py:427-428) is only consumed by the class-aware loss of the G step (trainer.py:438-446)
        x_imm = ops.nhwc_to_nchw_f32(x) if want_features else None
        cat = ops.concat_c(c, x)
        h, T["joint"] = self.joint.fwd(cat, training, groups=groups)
        T["h"] = h
        lg, ul = self.net.logits[0], self.net.uncond_logits[0]
        cond = ops.logits_fwd(h, lg.weight.detach(), lg.bias.detach(), out_cond)
        uncond = ops.logits_fwd(x, ul.weight.detach(), ul.bias.detach(), out_uncond)
        T["cond"], T["uncond"], T["c"] = cond, uncond, c
        return cond, uncond, x_imm, T

    def backward(self, T, dcond, duncond, dx_imm, need_dimg, need_dc, need_w=True, sink=None):
        """-> (grads dict or None when an external sink is used, dimg or None, dc or None)"""
        own = sink is None
        sink = GradSink() if own else sink
        groups = T.get("groups", 1)
        lg, ul = self.net.logits[0], self.net.uncond_logits[0]
        x, h = T["x_code"], T["h"]
        B = x.shape[0]
        dc = torch.zeros_like(T["c"])
        dx = None
        if need_w:   # `+=` kernels and never-reached branches need zeroed slots
            for p in (lg.weight, lg.bias, ul.weight, ul.bias):
                sink.zero_slot(p)
        if dcond is not None:
            dh = torch.empty_like(h)
            ops.logits_bwd(dcond, T["cond"], h, lg.weight.detach(), dh, False,
                           sink.g[lg.weight] if need_w else None, sink.g[lg.bias] if need_w else None)
            dcat = self.joint.bwd(T["joint"], dh, sink, need_w=need_w, groups=groups)
            dx = ops.concat_c_bwd(dcat, self.E, dc)
        elif need_w:
            for p in (self.joint.conv.weight, self.joint.bn.weight, self.joint.bn.bias):
                sink.zero_slot(p)
        if duncond is not None:
            first = dx is None
            if first:
                dx = torch.empty_like(x)
            ops.logits_bwd(duncond, T["uncond"], x, ul.weight.detach(), dx, not first,
                           sink.g[ul.weight] if need_w else None, sink.g[ul.bias] if need_w else None)
        if dx_imm is not None:
            _, H, W, C = x.shape
            dxi = ops.nchw_f32_to_nhwc(dx_imm, B, H, W, C, f32=self.precise)
            dx = dxi if dx is None else ops.add_bf16(dx, dxi)
        if dx is None:
            raise RuntimeError("sg2b200: D backward without any output gradient")
        for li in range(len(self.trunk) - 1, -1, -1):
            # the first trunk conv's dgrad also applies the stem's LeakyReLU backward (mask by the stem output)
            epi = (T["stem"][1], ops.EPI_LRELU_MASK) if li == 0 else None
            dx = self.trunk[li].bwd(T["trunk"][li], dx, sink, need_w=need_w, groups=groups, epi=epi)
        dimg = self.stem.bwd(T["stem"], dx, sink, need_dimg, need_w=need_w, masked=True)
        return (sink.finish() if own else None), dimg, (dc if need_dc else None)
