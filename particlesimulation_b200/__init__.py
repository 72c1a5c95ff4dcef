"""B200-native P3M / PM force step behind the ParticleSimulation C++ API.

The product is ``libp3m_b200.so`` (hand-written sm_100a CUDA kernels + cuFFT, C ABI in
``include/p3m_b200.h``) and the API-compatible C++ classes under ``host/``.  This Python package
only carries the ctypes binding used by the tests and ``bench.py`` and the synthetic
initial-condition generators of the bench harness.  Importing it never touches ``oracle/``.
"""
from . import capi  # noqa: F401
from .capi import Context, P3MError, P3MParams, default_params  # noqa: F401

__all__ = ["capi", "Context", "P3MError", "P3MParams", "default_params"]
