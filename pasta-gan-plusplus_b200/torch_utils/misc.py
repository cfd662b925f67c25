"""The three helpers of the reference's torch_utils/misc.py that touch the hot path
(assert_shape :86-99, profiled_function :104-109, suppress_tracer_warnings :62-68)."""
import contextlib
import functools
import warnings

import torch


@contextlib.contextmanager
def suppress_tracer_warnings():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=torch.jit.TracerWarning)
        yield


def assert_shape(tensor, ref_shape):
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref_size) in enumerate(zip(tensor.shape, ref_shape)):
        if ref_size is None:
            continue
        if int(size) != int(ref_size):
            raise AssertionError(f'Wrong size for dimension {idx}: got {size}, expected {ref_size}')


def profiled_function(fn):
    """Same profiler range names as the reference so traces line up."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    return wrapper
