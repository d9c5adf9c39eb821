"""Forward pass that substitutes test values for draws.

Same role as the reference's ``MetaSamplingExecutor``
(pymc4/flow/meta_executor.py:24-97): it discovers every variable's name, core
shape and kind without touching a random number generator; the samplers use it
to build the initial state (all chains start from the test point).
"""
from .transformed_executor import TransformedSamplingExecutor

__all__ = ["MetaSamplingExecutor"]


class MetaSamplingExecutor(TransformedSamplingExecutor):
    """Like the transformed executor, but 'sampling' returns ``dist.get_test_sample``."""

    def _draw(self, dist, sample_shape):
        if dist.is_root:
            return dist.get_test_sample(sample_shape=sample_shape)
        return dist.get_test_sample()
