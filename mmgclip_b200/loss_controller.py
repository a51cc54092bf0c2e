"""Name -> class lookup for losses (reference: mmgclip/loss/loss_controller.py:3-23).

``config.loss.config.loss_name`` selects the class, which the experiment then builds with no arguments
(ClassifierExperiment.py:70).
"""
from .losses import AveragedMedicalCLIPLoss, CLIPLoss, MMGCLIPLoss  # noqa: F401


def create_loss(loss_name):
    """Return the loss class called ``loss_name``; unknown names raise ``ValueError``."""
    loss_class = globals().get(loss_name, None)
    if loss_class is None or not isinstance(loss_class, type):
        raise ValueError(f"Invalid network_name: {loss_name}")
    return loss_class
