"""keras.tree subset."""


def is_nested(x):
    return isinstance(x, (list, tuple, dict))
