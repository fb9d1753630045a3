def mimsave(*a, **k):
    pass
