"""AMR error indicators and refinement by indicator (reference ``src/amr``).

The jump reductions run on the GPU, on the slab-order solution vector the solver
leaves there (``rtb200_err_angular_jump``); thresholds, hp steering and the mesh
refinement itself stay on the host, as in the reference.
"""
from .error_indicator import (Error_Indicator, Error_Indicator_Column,  # noqa: F401
                              Error_Indicator_Cell, from_file)
