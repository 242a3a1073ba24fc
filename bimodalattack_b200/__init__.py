"""B200-native drop-in for the per-step GCG+PGD hot path of Bimo99B9/BimodalAttack.

Same entry points as the reference package (``bimodalattack/__init__.py:1``)."""
from .config import BimodalAttackConfig, BimodalAttackResult
from .attack import BimodalAttack, run

__all__ = ["BimodalAttackConfig", "BimodalAttackResult", "BimodalAttack", "run"]
