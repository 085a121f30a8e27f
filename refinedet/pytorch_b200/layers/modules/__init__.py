from .refinedet_multibox_loss import RefineDetMultiBoxLoss

__all__ = ['RefineDetMultiBoxLoss']
