"""mmgclip_b200 -- B200-native (sm_100a) contrastive hot path of abdel-habib/mmg-clip.

Same-named drop-ins for the reference's projection heads (mmgclip/networks/projection.py), losses
(mmgclip/loss/losses.py), their name->class controllers, the MMGCLIP model shell and PromptClassifier
(mmgclip/networks/mmgclip_model.py), backed by hand-written CUDA kernels behind a C ABI (include/mmgclip_b200.h).
"""
from . import _lib  # noqa: F401  (does not load the shared object until first use)

__version__ = "0.1.0"
