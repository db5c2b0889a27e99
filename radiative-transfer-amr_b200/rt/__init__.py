"""RT solver layer: ``Problem`` (solve + operator entry points), the device
operator and the element tables (reference ``src/rt``)."""
from . import tables  # noqa: F401
from .operator import RTOperator  # noqa: F401
from .problem import Problem, TermOperator  # noqa: F401
