"""Numeric constants (reference: ``src/consts/__init__.py:9-16``)."""
import numpy as np

INT = np.int64
REAL = np.float64

PI = float(np.pi)
EPS = float(np.finfo(REAL).resolution)   # 1e-15, used as KSP rtol/atol
INF = float(np.finfo(REAL).max)

COMM_ROOT = 0
