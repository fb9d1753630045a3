from .mappo import MAPPOPolicy, MAPPOTrainer  # noqa: F401
from .valuenorm import ValueNorm  # noqa: F401
