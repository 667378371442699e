"""util/timing.py:6-11 of the reference: the duration logger the resampler wraps its phases in."""
import contextlib
import logging
import time


@contextlib.contextmanager
def log_duration(operation):
    logging.info(operation)
    start_time = time.time()
    yield
    logging.debug(f"{operation} took {time.time() - start_time:.2f} seconds")
