def init(*a, **k):
    pass


def log(*a, **k):
    pass


def finish(*a, **k):
    pass
