"""TEST INFRASTRUCTURE."""


def stochastic_rounding(value, dtype, seed=0):
    return value.to(dtype)
