"""Public dataclasses of the drop-in API. Field names, order and defaults mirror the reference's
``BimodalAttackConfig`` / ``BimodalAttackResult`` (bimodal_attack.py:42-85) -- including the fields
the reference never reads -- so ``BimodalAttackConfig(**kwargs)`` from its driver
(experiments.py:84-100) keeps working."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Union


@dataclass
class BimodalAttackConfig:
    num_steps: int = 250
    optim_str_init: Union[str, List[str]] = "x x x x x x x x x x x x x x x x x x x"
    search_width: int = 512
    batch_size: Optional[int] = None
    topk: int = 256
    n_replace: int = 1
    buffer_size: int = 0
    use_mellowmax: bool = False          # unused by the reference as well
    mellowmax_alpha: float = 1.0         # unused by the reference as well
    early_stop: bool = False
    allow_non_ascii: bool = False
    filter_ids: bool = True
    add_space_before_target: bool = False  # unused by the reference as well
    seed: Optional[int] = None
    verbosity: str = "INFO"
    dynamic_search: bool = False
    min_search_width: int = 32
    alpha: float = 0.01
    eps: float = 0.1
    pgd_attack: bool = False
    gcg_attack: bool = True
    debug_output: bool = False
    joint_eval: bool = False
    experiment_folder: str = "experiments/missing_folder"   # unused by the reference as well
    images_folder: str = "experiments/missing_folder/images"
    pgd_after_gcg: bool = False
    model: str = "llava"                 # unused by the reference as well


@dataclass
class BimodalAttackResult:
    best_loss: float
    best_string: str
    losses: List[float]
    strings: List[str]
    adversarial_suffixes: List[str]
    model_outputs: List[str]
    gradient_times: List[float]
    sampling_times: List[float]
    loss_times: List[float]
    pgd_times: List[float]
    total_times: List[float] = None
