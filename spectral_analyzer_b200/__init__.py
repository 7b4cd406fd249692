"""spectral_analyzer_b200 -- B200-native engine for the spectrogram / Welch-PSD / downconvert hot
path of GassiusODude/spectral_analyzer.  The product is libsa_engine.so (hand-written sm_100a
CUDA behind the C-ABI of include/sa_engine.h); this package is the Python mirror of the
reference's service interface on top of it.
"""
from .services import (Engine, EngineError, SpectralService, ExtractDownConvertService,  # noqa: F401
                       AsyncExtractDownConvertService, PowerSpectralDensity, IqData)

from .tiles import CanvasTileCache  # noqa: F401

__version__ = "0.1.0"
