"""Truely visual-analysis hot path, B200-native (sm_100a CUDA behind a C ABI).

Drop-in for reference ``server/model.py::run`` (server/model.py:11-95): the Python host code
in :mod:`model` keeps the reference's signature and return structure and drives
``libtruely_b200.so`` (``csrc/``, declared in ``include/truely_b200.h``) through ctypes.
PyTorch is used for device memory, streams and ``torch.distributed`` only.
There is no CPU fallback: every compute entry point raises if the CUDA library is absent.
"""
from . import synth, weights, _lib, model, service  # noqa: F401
from .model import run, run_trace, Analyzer  # noqa: F401

__version__ = "0.1.0"
