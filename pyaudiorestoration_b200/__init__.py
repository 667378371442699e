"""pyaudiorestoration_b200 -- B200-native STFT + varispeed-resample hot path of
HENDRIX-ZT2/pyaudiorestoration behind the reference's own module surface.

    from pyaudiorestoration_b200.util import fourier, resampling
    S = fourier.stft(x, n_fft=4096, step=1024)          # util/fourier.py:37
    resampling.run(files, speed_curve=curve, resampling_mode="Sinc", ...)   # util/resampling.py:162

All compute runs in hand-written sm_100a CUDA kernels inside ``libpar_b200.so`` (C ABI:
``include/par_b200.h``).  There is no CPU fallback: importing works anywhere, but any compute
call raises ``RuntimeError`` when the library or a CUDA device is missing.
"""
from . import _lib  # noqa: F401
from ._lib import build, library_path  # noqa: F401

__all__ = ["build", "library_path"]
__version__ = "0.1.0"
