import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)
        self.dtype = np.dtype(dtype)


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


class Tuple:
    def __init__(self, spaces):
        self.spaces = spaces


from . import box  # noqa: E402,F401
