"""Minimal stand-in for the `gym` package (absent from this image).

Test scaffolding only: lets the UNMODIFIED reference under /root/reference import.
The reference reads only `.shape/.n/.low/.high/.dtype` and the class *names*
`Box`/`Discrete` (it dispatches on `space.__class__.__name__`).
"""


class Env:
    metadata = {}

    def close(self):
        pass


class Space:
    def __init__(self, *a, **k):
        pass


from . import spaces, envs  # noqa: E402,F401
