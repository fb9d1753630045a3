"""Minimal gym-like spaces.  The reference only reads `.shape`, `.n`, `.low/.high` and dispatches on the
class NAME ("Box" / "Discrete"): algos/algo_utils/act.py:19-25, buffer/shared_buffer.py:51, utils/util.py:46-66."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box%s" % (self.shape,)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
