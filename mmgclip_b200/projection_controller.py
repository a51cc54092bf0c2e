"""Name -> class lookup for projection heads (reference: mmgclip/networks/projection_controller.py:3-24).

``config.projection.config.projection_name`` selects the class; ``"ZeroProjection"`` is handled by the caller
(mmgclip_model.py:36,47-49) and is *not* a valid name here, exactly as in the reference.
"""
from .projection import LinearProjectionLayer, MLPProjectionHead, MultiLinearHead  # noqa: F401


def get_projection_head(projection_name):
    """Return the projection-head class called ``projection_name``; unknown names raise ``ValueError``."""
    head_class = globals().get(projection_name, None)
    if head_class is None or not isinstance(head_class, type):
        raise ValueError(f"Invalid network_name: {projection_name}")
    return head_class
