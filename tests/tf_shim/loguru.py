"""Import stub (tests/tf_shim): lets the reference modules import where this third-party package is absent. TEST INFRASTRUCTURE."""


class _Logger(object):
    def __getattr__(self, name):
        return lambda *a, **k: None


logger = _Logger()
