"""Import stub (tests/tf_shim): lets the reference modules import where this third-party package is absent. TEST INFRASTRUCTURE."""


def compress(*a, **k):
    raise NotImplementedError


def decompress(*a, **k):
    raise NotImplementedError
