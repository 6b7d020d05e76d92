from .advantages import GAE
from .functional import generalized_advantage_estimate, vec_generalized_advantage_estimate

__all__ = ["GAE", "generalized_advantage_estimate", "vec_generalized_advantage_estimate"]
