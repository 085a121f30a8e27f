from .refinedet_multibox_loss import RefineDetMultiBoxLoss, RefineDetCriterionPair

__all__ = ['RefineDetMultiBoxLoss', 'RefineDetCriterionPair']
