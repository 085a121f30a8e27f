"""Drop-in for ``layers/modules/refinedet_multibox_loss.py`` (reference :10-139).

The per-image ``refine_match`` Python loop (:75-86) becomes one batched kernel pair and the
double sort of the hard-negative mining (:119-123) becomes one radix-select kernel; SmoothL1,
``log_sum_exp`` and cross-entropy stay on stock PyTorch because they need autograd
(SURVEY.md §2, §8a a11).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..box_utils import LABEL_ARM_BINARY, LABEL_ODM, hnm_select, log_sum_exp, match_batch, pad_targets

# data/config.py:57 (``coco['variance']``, read at reference :46)
_VARIANCE = [0.1, 0.2]


class RefineDetMultiBoxLoss(nn.Module):
    """RefineDet weighted loss: SmoothL1 localisation + cross-entropy confidence with 3:1 hard
    negative mining, for the ARM (``use_ARM=False``, 2-class) or ODM (``use_ARM=True``)
    branch.  Constructor arguments are the reference's (:33-48)."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching, bkg_label, neg_mining,
                 neg_pos, neg_overlap, encode_target, use_gpu=True, theta=0.01, use_ARM=False):
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

    def match_targets(self, predictions, targets):
        """Targets of reference :62-90 for the whole batch: ``(loc_t[B,P,4], conf_t[B,P])``."""
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        loc_data = odm_loc_data if self.use_ARM else arm_loc_data
        if not loc_data.is_cuda:
            raise RuntimeError('RefineDetMultiBoxLoss: predictions must be CUDA tensors '
                               '(refinedet.pytorch_b200 has no CPU fallback)')
        for t in targets:
            if t.shape[0] == 0:
                raise IndexError('RefineDetMultiBoxLoss: an image has no ground-truth boxes '
                                 '(the reference raises in refine_match, box_utils.py:139)')
        priors = priors[:loc_data.size(1), :]                       # :68 (DataParallel gather)
        truths, labels, gt_count = pad_targets(targets, loc_data.device)
        if self.num_classes == 2 and not self.use_ARM:
            mode = LABEL_ARM_BINARY                                 # labels = labels >= 0, :78-79
        else:
            mode = LABEL_ODM
        arm = arm_loc_data.detach() if self.use_ARM else None       # :80-85
        return match_batch(self.threshold, truths, labels, gt_count, priors.detach(), self.variance, arm, mode)

    def forward(self, predictions, targets):
        """reference :50-139.  ``predictions`` = (arm_loc, arm_conf, odm_loc, odm_conf, priors),
        ``targets`` = list of ``[G_i,5]`` tensors.  Returns ``(loss_l, loss_c)``."""
        arm_loc_data, arm_conf_data, odm_loc_data, odm_conf_data, priors = predictions
        if self.use_ARM:
            loc_data, conf_data = odm_loc_data, odm_conf_data
        else:
            loc_data, conf_data = arm_loc_data, arm_conf_data
        num = loc_data.size(0)
        loc_t, conf_t = self.match_targets(predictions, targets)

        pos = conf_t > 0
        if self.use_ARM:                                            # :96-101
            P = F.softmax(arm_conf_data, 2)
            pos = pos & ~(P[:, :, 1] <= self.theta).detach()

        # Localization Loss (Smooth L1), :105-110
        pos_idx = pos.unsqueeze(pos.dim()).expand_as(loc_data)
        loc_p = loc_data[pos_idx].view(-1, 4)
        loc_tp = loc_t[pos_idx].view(-1, 4)
        loss_l = F.smooth_l1_loss(loc_p, loc_tp, reduction='sum')

        # per-anchor confidence loss for mining, :113-114
        batch_conf = conf_data.view(-1, self.num_classes)
        loss_c = log_sum_exp(batch_conf) - batch_conf.gather(1, conf_t.view(-1, 1))

        # Hard Negative Mining, :117-123 (positives are zeroed inside the kernel)
        neg, num_pos = hnm_select(loss_c.detach().view(num, -1), pos, self.negpos_ratio)

        # Confidence Loss Including Positive and Negative Examples, :126-130
        sel = pos | neg
        conf_p = conf_data[sel.unsqueeze(2).expand_as(conf_data)].view(-1, self.num_classes)
        targets_weighted = conf_t[sel]
        loss_c = F.cross_entropy(conf_p, targets_weighted, reduction='sum')

        N = num_pos.sum().float()                                   # :134
        if N < 1:                                                   # :135-136
            return torch.zeros(1), torch.zeros(1)
        loss_l = loss_l / N
        loss_c = loss_c / N
        return loss_l, loss_c
