"""B200-native (sm_100a) implementation of Move2Hear's per-step audio-separation hot path.

RIR convolution + quantise + mix + STFT (``render``) and the passive-separator U-Net (``separator_cnn``),
behind the reference's own Python call signatures, backed by the C-ABI library ``libm2h_b200.so``
(``include/m2h.h``).  No CPU fallback, no Triton, no multi-backend dispatch.
"""
from ._lib import M2HError, build, load  # noqa: F401

__version__ = "0.1.0"
