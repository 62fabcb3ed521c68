"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle.c).  Parity status:
unpinned for torch_scatter/torch_sparse entry points (upstream absent),
pinned by tests/golden for the native-torch ops the reference calls."""
from .oracle import *  # noqa: F401,F403
