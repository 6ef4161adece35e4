"""mcmc_b200 — B200-native many-chain HMC / MALA / NUTS / RM-HMC (+ RWMH) engine.

The product is ``libmcmc_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/mcmc_b200.h``) plus the C++ drop-in header ``include/mcmc_b200.hpp``.
``mcmc_b200.api`` is a thin ctypes binding used by the tests and ``bench.py``.
"""
from . import api  # noqa: F401
from .api import McmcB200Error, de, hmc, mala, nuts, rmhmc, rwmh  # noqa: F401
