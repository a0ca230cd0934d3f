"""adt_b200: B200-native hot path of defineZYP/ADT (SASRec-ADT training step + full-catalog evaluation)."""
from .model import SASRecADT  # noqa: F401
from .lambdas import get_lambdas, get_weight  # noqa: F401

__all__ = ["SASRecADT", "get_lambdas", "get_weight"]
