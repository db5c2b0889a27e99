"""DG substrate: hp mesh, quadrature, nodal projections, DoF numbering and the
flattened device layout (reference ``src/dg``)."""
from . import mesh, quadrature, projection, matrix  # noqa: F401
from .mesh import Mesh, Column, Cell  # noqa: F401
from .projection import Projection, push_forward, pull_back  # noqa: F401
from .quadrature import quad_xyth, lag_eval, lag_ddx_eval  # noqa: F401
from .matrix import (get_idx_map, get_idx_inv_map, get_col_idxs, get_cell_idxs,  # noqa: F401
                     get_intr_mask, get_bdry_mask, split_vector, merge_vectors)
