"""prospect_public_b200 -- B200 (sm_100a) implementation of PROSPECT's simulated-annealing /
Metropolis-Hastings hot path over the analytic Gaussian kernels.

Host code (this package) mirrors the reference's interfaces for that path; all numerics run in
hand-written CUDA kernels behind the C ABI of include/pb200.h (libpb200.so).  See DESIGN.md.
"""
__version__ = '0.1.0'

from .config import Configuration, ConfigError  # noqa: F401
from .io import load_profile  # noqa: F401
