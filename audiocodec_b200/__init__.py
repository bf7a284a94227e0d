"""audiocodec_b200: the encode/decode hot path of korneelvdbroek/audiocodec as sm_100a CUDA kernels.

    from audiocodec_b200.mdctransformer import MDCTransformer
    from audiocodec_b200.psychoacoustic import PsychoacousticModel

mirror `audiocodec.mdctransformer` / `audiocodec.psychoacoustic` of the reference.  The classes call a
C-ABI shared library (include/audiocodec_b200.h) through ctypes; there is no CPU fallback.
"""

from .mdctransformer import MDCTransformer
from .psychoacoustic import PsychoacousticModel
from .codec import AudioCodec

__all__ = ["MDCTransformer", "PsychoacousticModel", "AudioCodec"]
__version__ = "0.1.0"
