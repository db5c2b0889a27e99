"""rt_b200 -- B200-native transport-operator path for hp-adaptive DGSEM
radiative transfer (2-D space x 1-D angle).

Sub-packages mirror the reference layout (``/root/reference/src``):

* ``consts``  -- INT/REAL/PI/EPS/INF            (reference ``src/consts``)
* ``dg``      -- hp mesh, quadrature, projection, index maps, FlatMesh layout
                 (reference ``src/dg``)
* ``rt``      -- ``Problem`` with ``solve`` and the operator entry points
                 (reference ``src/rt``), running on the CUDA library
                 ``csrc/librtb200.so`` through ctypes.

There is no CPU fallback: operator applies and the Krylov solve fail loudly
when the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
