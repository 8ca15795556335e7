"""Import stub (tests/tf_shim): lets the reference modules import where this third-party package is absent. TEST INFRASTRUCTURE."""


def display(*a, **k):
    pass


def HTML(*a, **k):
    return None
