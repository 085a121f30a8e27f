"""Drop-in for ``layers/modules/refinedet_multibox_loss.py`` (reference :10-139).

The per-image ``refine_match`` Python loop (:75-86) becomes one batched kernel pair, the double
sort of the hard-negative mining (:119-123) one radix-select kernel, and the loss tail
(:96-138: ARM-theta gate, ``log_sum_exp - gather``, SmoothL1, cross-entropy, ``/ N``) three
kernels forward and one backward behind a ``torch.autograd.Function`` — ``conf_data`` is read once
forward and only on the ``pos | neg`` rows backward.
"""
import ctypes
import threading

import torch
import torch.nn as nn

from ctypes import c_void_p

from ..._ffi import check, lib, on_device, ptr, require_cuda_f32, stream_ptr
from ..box_utils import LABEL_ARM_BINARY, LABEL_ODM, match_batch, _padded

# data/config.py:57 (``coco['variance']``, read at reference :46)
_VARIANCE = [0.1, 0.2]


class RefineDetMultiBoxLoss(nn.Module):
    """RefineDet weighted loss: SmoothL1 localisation + cross-entropy confidence with 3:1 hard
    negative mining, for the ARM (``use_ARM=False``, 2-class) or ODM (``use_ARM=True``)
    branch.  Constructor arguments are the reference's (:33-48).

    ``sync_free`` (extension, default off): the reference reads ``N = sum(num_pos)`` on the host to return
    ``(zeros(1), zeros(1))`` when no positive survives (:135-136) — a stream drain in the middle of every
    criterion call.  With ``sync_free=True`` nothing is read back: the losses are always the 0-dim device
    tensors, and for ``N < 1`` the kernels deliver exactly zero losses and zero gradients — the same parameter
    update as the reference's no-grad zeros, so ``train_refinedet.py:252-261`` runs unchanged while the host
    queues the next criterion and the backward pass behind the running kernels."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining,
                 neg_pos, neg_overlap, encode_target, use_gpu=True, theta=0.01, use_ARM=False, sync_free=False):
        super(RefineDetMultiBoxLoss, self).__init__()
        self.use_gpu = use_gpu
        self.num_classes = num_classes
        self.threshold = overlap_thresh
        self.background_label = bkg_label
        self.encode_target = encode_target
        self.use_prior_for_matching = prior_for_matching
        self.do_neg_mining = neg_mining
        self.negpos_ratio = neg_pos
        self.neg_overlap = neg_overlap
        self.variance = _VARIANCE
        self.theta = theta
        self.use_ARM = use_ARM
        self.sync_free = bool(sync_free)

    def match_targets(self, predictions, targets):
        """Targets of reference :62-90 for the whole batch: ``(loc_t[B,P,4], conf_t[B,P])``."""
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        loc_data = odm_loc_data if self.use_ARM else arm_loc_data
        if not loc_data.is_cuda:
            raise RuntimeError('RefineDetMultiBoxLoss: predictions must be CUDA tensors '
                               '(refinedet.pytorch_b200 has no CPU fallback)')
        truths, labels, gt_count, min_count = _padded(targets, loc_data.device)
        if min_count == 0:
            raise IndexError('RefineDetMultiBoxLoss: an image has no ground-truth boxes '
                             '(the reference raises in refine_match, box_utils.py:139)')
        priors = priors[:loc_data.size(1), :]                       # :68 (DataParallel gather)
        if self.num_classes == 2 and not self.use_ARM:
            mode = LABEL_ARM_BINARY                                 # labels = labels >= 0, :78-79
        else:
            mode = LABEL_ODM
        arm = arm_loc_data.detach() if self.use_ARM else None       # :80-85
        return match_batch(self.threshold, truths, labels, gt_count, priors.detach(), self.variance, arm, mode)

    def forward(self, predictions, targets):
        """reference :50-139.  ``predictions`` = (arm_loc, arm_conf, odm_loc, odm_conf, priors),
        ``targets`` = list of ``[G_i,5]`` tensors.  Returns ``(loss_l, loss_c)``.

        One native call (``rd_multibox_criterion``: match, confidence loss, mining, reductions — six kernels chained
        on the device) behind a ``torch.autograd.Function``; backward is one more (``rd_multibox_loss_backward``)."""
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        if self.use_ARM:
            loc_data, conf_data = odm_loc_data, odm_conf_data
        else:
            loc_data, conf_data = arm_loc_data, arm_conf_data
        if not loc_data.is_cuda:
            raise RuntimeError('RefineDetMultiBoxLoss: predictions must be CUDA tensors '
                               '(refinedet.pytorch_b200 has no CPU fallback)')
        truths, labels, gt_count, min_count = _padded(targets, loc_data.device)
        if min_count == 0:
            raise IndexError('RefineDetMultiBoxLoss: an image has no ground-truth boxes '
                             '(the reference raises in refine_match, box_utils.py:139)')
        priors = priors[:loc_data.size(1), :]                                 # :68 (DataParallel gather)
        mode = LABEL_ARM_BINARY if (self.num_classes == 2 and not self.use_ARM) else LABEL_ODM     # :78-79
        arm_loc = arm_loc_data.detach() if self.use_ARM else None             # :80-85
        arm_gate = arm_conf_data.detach() if self.use_ARM else None           # :96-101 (softmax inside the kernel)
        state = _CriterionState()
        loss_l, loss_c, N = _Criterion.apply(loc_data, conf_data, state, truths, labels, gt_count, priors.detach(),
                                             arm_loc, arm_gate, float(self.threshold), self.variance, mode,
                                             float(self.theta), int(self.negpos_ratio))
        self._last = state
        if not self.sync_free and float(N) < 1:                     # :135-136 (the reference syncs here too)
            return torch.zeros(1), torch.zeros(1)
        return loss_l, loss_c

    @property
    def last_masks(self):
        """``(pos, neg)`` bool ``[B,P]`` of this criterion's latest forward (inspection / tests)."""
        st = self._last
        return st.view('pos', torch.bool), st.view('neg', torch.bool)

    @property
    def last_targets(self):
        """``(loc_t, conf_t)`` of this criterion's latest forward."""
        st = self._last
        return st.view('loc_t', torch.float32, 4), st.view('conf_t', torch.int64)


def _up(n):
    return (n + 255) // 256 * 256


class _CriterionState(object):
    """Everything one criterion forward leaves on the device for its backward, in ONE allocation:
    ``loc_t | conf_t | ce | lse | pos | neg | num_pos | workspace`` (256-byte aligned regions)."""
    __slots__ = ('buf', 'off', 'B', 'P', 'losses')
    _FIELDS = (('loc_t', 16), ('conf_t', 8), ('ce', 4), ('lse', 4), ('pos', 1), ('neg', 1))

    def allocate(self, B, P, G, device):
        self.B, self.P = B, P
        off, o = {}, 0
        for name, width in self._FIELDS:
            off[name] = o
            o += _up(B * P * width)
        off['num_pos'] = o
        o += _up(B * 4)
        off['ws'] = o
        ws_bytes = int(lib().rd_multibox_criterion_workspace_bytes(B, P, G))
        self.off = off
        self.buf = torch.empty(o + ws_bytes, dtype=torch.uint8, device=device)
        self.losses = torch.empty(3, dtype=torch.float32, device=device)
        return ws_bytes

    def ptr(self, name):
        return c_void_p(self.buf.data_ptr() + self.off[name])

    def view(self, name, dtype, inner=None):
        width = dict(self._FIELDS)[name]
        t = self.buf[self.off[name]:self.off[name] + self.B * self.P * width].view(dtype)
        return t.view(self.B, self.P, inner) if inner else t.view(self.B, self.P)


class _Criterion(torch.autograd.Function):
    """refinedet_multibox_loss.py:62-138 on the device: forward = rd_multibox_criterion, backward =
    rd_multibox_loss_backward.  Differentiable in ``loc_data`` and ``conf_data`` only."""

    @staticmethod
    def forward(ctx, loc_data, conf_data, state, truths, labels, gt_count, priors, arm_loc, arm_gate, threshold,
                variance, mode, theta, negpos_ratio):
        loc_c = require_cuda_f32(loc_data, 'loc_data')
        conf_c = require_cuda_f32(conf_data, 'conf_data', align=8)
        priors = require_cuda_f32(priors, 'priors')
        if arm_loc is not None:
            arm_loc = require_cuda_f32(arm_loc, 'arm_loc_data')
        if arm_gate is not None:
            arm_gate = require_cuda_f32(arm_gate, 'arm_conf_data', align=8)
        B, P = loc_c.shape[0], loc_c.shape[1]
        C = conf_c.shape[-1]
        G = truths.shape[1]
        if conf_c.numel() != B * P * C or priors.shape[0] != P:
            raise ValueError('conf_data must be [B,P,C] and priors [P,4] for loc_data [B,P,4]')
        ws_bytes = state.allocate(B, P, G, loc_c.device)
        with on_device(loc_c.device):
            check(lib().rd_multibox_criterion(
                ptr(truths), ptr(labels), ptr(gt_count), ptr(priors), ptr(arm_loc), ptr(loc_c), ptr(conf_c), ptr(arm_gate),
                B, P, C, G, threshold, float(variance[0]), float(variance[1]), mode, theta, negpos_ratio,
                state.ptr('ws'), ws_bytes, state.ptr('loc_t'), state.ptr('conf_t'), state.ptr('ce'), state.ptr('lse'),
                state.ptr('pos'), state.ptr('neg'), state.ptr('num_pos'), ptr(state.losses), stream_ptr()),
                'rd_multibox_criterion')
        loss_l, loss_c, N = state.losses.unbind(0)
        ctx.save_for_backward(loc_c, conf_c)
        ctx.state, ctx.keep = state, (truths, labels, gt_count, priors, arm_loc, arm_gate)
        ctx.mark_non_differentiable(N)
        return loss_l, loss_c, N

    @staticmethod
    def backward(ctx, g_l, g_c, _g_n):
        loc_c, conf_c = ctx.saved_tensors
        st = ctx.state
        need_loc, need_conf = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_l = g_l.contiguous().float() if g_l is not None else None
        g_c = g_c.contiguous().float() if g_c is not None else None
        grad_loc = torch.empty_like(loc_c) if need_loc else None
        grad_conf = torch.empty_like(conf_c) if need_conf else None
        n_dev = c_void_p(st.losses.data_ptr() + 8)
        with on_device(conf_c.device):
            check(lib().rd_multibox_loss_backward(ptr(loc_c), st.ptr('loc_t'), ptr(conf_c), st.ptr('conf_t'), st.ptr('lse'),
                                                  st.ptr('pos'), st.ptr('neg'), ptr(g_l), ptr(g_c), n_dev,
                                                  st.B * st.P, conf_c.shape[-1], ptr(grad_loc), ptr(grad_conf), stream_ptr()),
                  'rd_multibox_loss_backward')
        return (grad_loc, grad_conf) + (None,) * 12


# ---------------------------------------------------------------------------------------------------------------
# both criteria of a training step in one call
# ---------------------------------------------------------------------------------------------------------------
_tls = threading.local()
_STATE_FIELDS = ('loc_t', 'conf_t', 'ce', 'lse', 'pos', 'neg', 'num_pos', 'losses', 'ws')
_layouts = {}


def _state_layout(B, P, G):
    """``(state_bytes, {field: offset})`` of one criterion state (``rd_criterion_state_layout``), cached per shape."""
    key = (B, P, G)
    lay = _layouts.get(key)
    if lay is None:
        offs = (ctypes.c_size_t * 9)()
        check(lib().rd_criterion_state_layout(B, P, G, offs), 'rd_criterion_state_layout')
        lay = _layouts[key] = (int(lib().rd_criterion_state_bytes(B, P, G)), dict(zip(_STATE_FIELDS, [int(o) for o in offs])))
    return lay


def _side_stream(device):
    """This thread's side stream on ``device`` (the ODM chain of a pair runs there, forked from the current stream)."""
    streams = getattr(_tls, 'side', None)
    if streams is None:
        streams = _tls.side = {}
    st = streams.get(device.index)
    if st is None:
        st = streams[device.index] = torch.cuda.Stream(device=device)
    return st


class _PairState(object):
    """The states of the ARM and the ODM criterion of one step: ONE device allocation ``[arm state | odm state]``."""
    __slots__ = ('buf', 'off', 'nbytes', 'B', 'P', 'G')
    _WIDTH = {'loc_t': 16, 'conf_t': 8, 'ce': 4, 'lse': 4, 'pos': 1, 'neg': 1}

    def allocate(self, B, P, G, device):
        self.B, self.P, self.G = B, P, G
        self.nbytes, self.off = _state_layout(B, P, G)
        self.buf = torch.empty(2 * self.nbytes, dtype=torch.uint8, device=device)

    def base(self, which):
        return self.buf.data_ptr() + which * self.nbytes

    def losses(self, which):
        o = which * self.nbytes + self.off['losses']
        return self.buf[o:o + 12].view(torch.float32)

    def view(self, which, name, dtype, inner=None):
        o = which * self.nbytes + self.off[name]
        t = self.buf[o:o + self.B * self.P * self._WIDTH[name]].view(dtype)
        return t.view(self.B, self.P, inner) if inner else t.view(self.B, self.P)


class _CriterionPair(torch.autograd.Function):
    """train_refinedet.py:252-253 on the device: forward = rd_multibox_criterion_pair (twelve kernels, the two chains
    on two streams), backward = rd_multibox_loss_backward_pair.  Differentiable in the four prediction tensors."""

    @staticmethod
    def forward(ctx, arm_loc, arm_conf, odm_loc, odm_conf, state, truths, labels, gt_count, priors, arm_threshold,
                odm_threshold, variance, arm_mode, theta, arm_negpos, odm_negpos, concurrent):
        arm_loc_c = require_cuda_f32(arm_loc, 'arm_loc_data')
        arm_conf_c = require_cuda_f32(arm_conf, 'arm_conf_data', align=8)
        odm_loc_c = require_cuda_f32(odm_loc, 'odm_loc_data')
        odm_conf_c = require_cuda_f32(odm_conf, 'odm_conf_data', align=8)
        priors = require_cuda_f32(priors, 'priors')
        B, P = odm_loc_c.shape[0], odm_loc_c.shape[1]
        C = odm_conf_c.shape[-1]
        G = truths.shape[1]
        if (arm_loc_c.numel() != B * P * 4 or arm_conf_c.numel() != B * P * 2 or odm_conf_c.numel() != B * P * C
                or priors.shape[0] != P):
            raise ValueError('predictions must be arm_loc [B,P,4], arm_conf [B,P,2], odm_loc [B,P,4], odm_conf [B,P,C], '
                             'priors [P,4]')
        dev = odm_loc_c.device
        state.allocate(B, P, G, dev)
        with on_device(dev):
            side = c_void_p(_side_stream(dev).cuda_stream) if concurrent else c_void_p(0)
            check(lib().rd_multibox_criterion_pair(
                ptr(truths), ptr(labels), ptr(gt_count), ptr(priors), ptr(arm_loc_c), ptr(arm_conf_c), ptr(odm_loc_c),
                ptr(odm_conf_c), B, P, C, G, arm_threshold, odm_threshold, float(variance[0]), float(variance[1]), arm_mode,
                theta, arm_negpos, odm_negpos, c_void_p(state.base(0)), c_void_p(state.base(1)), state.nbytes, stream_ptr(),
                side), 'rd_multibox_criterion_pair')
        al, ac, an = state.losses(0).unbind(0)
        ol, oc, on = state.losses(1).unbind(0)
        ctx.save_for_backward(arm_loc_c, arm_conf_c, odm_loc_c, odm_conf_c)
        ctx.state, ctx.keep, ctx.concurrent = state, (truths, labels, gt_count, priors), concurrent
        ctx.mark_non_differentiable(an, on)
        return al, ac, an, ol, oc, on

    @staticmethod
    def backward(ctx, g_al, g_ac, _g_an, g_ol, g_oc, _g_on):
        arm_loc_c, arm_conf_c, odm_loc_c, odm_conf_c = ctx.saved_tensors
        st = ctx.state
        gs = [g.contiguous().float() if g is not None else None for g in (g_al, g_ac, g_ol, g_oc)]
        outs = [torch.empty_like(t) if need else None
                for t, need in zip((arm_loc_c, arm_conf_c, odm_loc_c, odm_conf_c), ctx.needs_input_grad[:4])]
        dev = odm_conf_c.device
        with on_device(dev):
            side = c_void_p(_side_stream(dev).cuda_stream) if ctx.concurrent else c_void_p(0)
            check(lib().rd_multibox_loss_backward_pair(
                ptr(arm_loc_c), ptr(arm_conf_c), ptr(odm_loc_c), ptr(odm_conf_c), c_void_p(st.base(0)), c_void_p(st.base(1)),
                st.B, st.P, odm_conf_c.shape[-1], st.G, ptr(gs[0]), ptr(gs[1]), ptr(gs[2]), ptr(gs[3]), ptr(outs[0]),
                ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), stream_ptr(), side), 'rd_multibox_loss_backward_pair')
        return tuple(outs) + (None,) * 13


class RefineDetCriterionPair(nn.Module):
    """The ARM and the ODM criterion of a RefineDet training step in ONE call (extension; the reference calls its two
    ``RefineDetMultiBoxLoss`` modules one after the other on the same ``(predictions, targets)``,
    ``train_refinedet.py:252-253``):

        pair = RefineDetCriterionPair(arm_criterion, odm_criterion)
        arm_loss_l, arm_loss_c, odm_loss_l, odm_loss_c = pair(out, targets)

    One native call forward (``rd_multibox_criterion_pair``: targets padded once, the two kernel chains run
    concurrently on two streams) and one backward, one ``autograd.Function`` instead of two.  Values and gradients are
    those of the two modules called separately (the same kernels on the same inputs).  When either module has
    ``sync_free=False`` the two ``N`` are read after both chains were queued (one wait for the step instead of two in
    the middle of it) and a criterion with ``N < 1`` returns the reference's ``(zeros(1), zeros(1))`` (:135-136)."""

    def __init__(self, arm_criterion, odm_criterion, concurrent=True):
        super(RefineDetCriterionPair, self).__init__()
        if arm_criterion.use_ARM or not odm_criterion.use_ARM:
            raise ValueError('RefineDetCriterionPair(arm_criterion [use_ARM=False], odm_criterion [use_ARM=True])')
        if arm_criterion.num_classes != 2:
            raise ValueError('the ARM criterion of a pair is the binary one (num_classes=2, train_refinedet.py:124)')
        self.arm_criterion, self.odm_criterion = arm_criterion, odm_criterion
        self.concurrent = bool(concurrent)
        self.sync_free = arm_criterion.sync_free and odm_criterion.sync_free

    def forward(self, predictions, targets):
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        if not odm_loc_data.is_cuda:
            raise RuntimeError('RefineDetCriterionPair: predictions must be CUDA tensors '
                               '(refinedet.pytorch_b200 has no CPU fallback)')
        a, o = self.arm_criterion, self.odm_criterion
        if odm_conf_data.shape[-1] != o.num_classes:
            raise ValueError('odm_conf_data has %d classes, the ODM criterion %d' % (odm_conf_data.shape[-1], o.num_classes))
        truths, labels, gt_count, min_count = _padded(targets, odm_loc_data.device)
        if min_count == 0:
            raise IndexError('RefineDetMultiBoxLoss: an image has no ground-truth boxes '
                             '(the reference raises in refine_match, box_utils.py:139)')
        priors = priors[:odm_loc_data.size(1), :]                             # :68 (DataParallel gather)
        state = _PairState()
        al, ac, an, ol, oc, on = _CriterionPair.apply(
            arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, state, truths, labels, gt_count, priors.detach(),
            float(a.threshold), float(o.threshold), o.variance, LABEL_ARM_BINARY, float(o.theta), int(a.negpos_ratio),
            int(o.negpos_ratio), self.concurrent)
        a._last = _StateView(state, 0)
        o._last = _StateView(state, 1)
        if not self.sync_free:
            if float(an) < 1:                                       # :135-136 (the reference syncs here too)
                al, ac = torch.zeros(1), torch.zeros(1)
            if float(on) < 1:
                ol, oc = torch.zeros(1), torch.zeros(1)
        return al, ac, ol, oc


class _StateView(object):
    """One half of a :class:`_PairState` behind the ``view`` interface of :class:`_CriterionState`
    (``last_masks`` / ``last_targets`` of the two modules after a paired call)."""
    __slots__ = ('state', 'which')

    def __init__(self, state, which):
        self.state, self.which = state, which

    def view(self, name, dtype, inner=None):
        return self.state.view(self.which, name, dtype, inner)
