"""Model classes with the reference's constructor / method surface (src/models/__init__.py:1-7)."""
from .exact_gp import ExactGP
from .variational_gp import VariationalGP
from .exact_cmp import ExactCMP
from .variational_cmp import VariationalCMP
from .cmp import CMP
from .bagged_gp import BaggedGP

__all__ = ["ExactGP", "VariationalGP", "ExactCMP", "VariationalCMP", "CMP", "BaggedGP"]
