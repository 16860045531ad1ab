"""mrclip_b200 -- B200-native (sm_100a) distributed contrastive loss for MR-CLIP.

Public surface mirrors the reference's ``open_clip.loss`` for the hot path:
``ClipLoss``, ``SigLipLoss``, ``MultiPositiveClipLoss``, ``gather_features`` and the ``neighbour_exchange*`` ring hops.  The kernels live in ``csrc/`` behind the C ABI
declared in ``include/mrclip.h`` and are loaded with ctypes (``_cabi.py``).
"""
from .exchange import (NeighbourExchange, NeighbourExchangeBidir, neighbour_exchange,  # noqa: F401
                       neighbour_exchange_bidir, neighbour_exchange_bidir_with_grad, neighbour_exchange_with_grad)
from .loss import (ClipLoss, MultiPositiveClipLoss, SigLipLoss, gather_features, gather_features_with_tokens,  # noqa: F401
                   multi_positive_cross_entropy_loss, set_engine)

__version__ = "0.1.0"
__all__ = ["ClipLoss", "SigLipLoss", "MultiPositiveClipLoss", "gather_features", "gather_features_with_tokens",
           "multi_positive_cross_entropy_loss", "neighbour_exchange",
           "neighbour_exchange_bidir", "neighbour_exchange_with_grad", "neighbour_exchange_bidir_with_grad",
           "NeighbourExchange", "NeighbourExchangeBidir"]
