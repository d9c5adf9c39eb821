"""Model interpreters (executors) and the sampling state."""
from .executor import SamplingExecutor, SamplingState, EvaluationError, StopExecution
from .transformed_executor import TransformedSamplingExecutor
from .meta_executor import MetaSamplingExecutor

__all__ = [
    "SamplingExecutor",
    "SamplingState",
    "TransformedSamplingExecutor",
    "MetaSamplingExecutor",
    "evaluate_model",
    "evaluate_model_transformed",
    "evaluate_meta_model",
]

evaluate_model = SamplingExecutor()
evaluate_model_transformed = TransformedSamplingExecutor()
evaluate_meta_model = MetaSamplingExecutor()
