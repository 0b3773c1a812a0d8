"""TEST-ONLY: the CPU-oracle-backed native op used to drive the host layer (see oracle/cpu_backend.py)."""
from oracle.cpu_backend import PortTensorQuantizer as OracleTensorQuantizer  # noqa: F401
from oracle.cpu_backend import ReferenceTensorQuantizer  # noqa: F401
