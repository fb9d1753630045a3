from .make_env import make_env  # noqa: F401
from .cuda_vec_env import CudaVecEnv  # noqa: F401
