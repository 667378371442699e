"""Mirror of the reference's ``util`` package for the hot path: ``fourier``, ``resampling`` (+ the
``io_ops`` / ``timing`` helpers they import).  Modules are imported on attribute access so that a
CPU-only import of one of them does not pull the other."""
import importlib

__all__ = ["fourier", "resampling", "io_ops", "timing", "flac", "wow_detection"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
