"""B200-native batched powered-descent environment + PSO fitness evaluation."""
from .params import RocketParams  # noqa: F401
