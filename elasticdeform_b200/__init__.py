"""elasticdeform_b200 -- B200-native drop-in for the hot path of gvtulder/elasticdeform.

    import elasticdeform_b200 as elasticdeform
    Y = elasticdeform.deform_random_grid(X, sigma=25, points=3)

Same public names as the reference package (reference __init__.py:1).
"""
from .deform_grid import deform_random_grid, deform_grid, deform_grid_gradient

__version__ = "0.1.0"
__all__ = ["deform_random_grid", "deform_grid", "deform_grid_gradient"]
