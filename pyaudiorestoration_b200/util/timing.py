"""Phase timing for the resampler: the reference wraps "Preparing" / "Resampling" / "Writing" in
``log_duration`` (util/timing.py:6-11) and the GUIs' log windows show those lines, so the same
two records are emitted here: the phase name at INFO when it starts, its wall time at DEBUG when
it ends (also when the phase raises)."""
import logging
from time import perf_counter


class log_duration:
    """``with log_duration("Resampling"): ...`` -- usable as a context manager or a decorator."""

    def __init__(self, operation, logger=None):
        self.operation = operation
        self.logger = logger or logging.getLogger()
        self.seconds = None
        self._t0 = None

    def __enter__(self):
        self.logger.info(self.operation)
        self._t0 = perf_counter()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.seconds = perf_counter() - self._t0
        self.logger.debug("%s took %.2f seconds", self.operation, self.seconds)
        return False

    def __call__(self, fn):
        def wrapped(*args, **kwargs):
            with log_duration(self.operation, self.logger):
                return fn(*args, **kwargs)
        wrapped.__name__ = getattr(fn, "__name__", "wrapped")
        wrapped.__doc__ = fn.__doc__
        return wrapped
