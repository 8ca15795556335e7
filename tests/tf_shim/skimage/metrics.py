"""Import stub (tests/tf_shim): lets the reference modules import where this third-party package is absent. TEST INFRASTRUCTURE."""


def structural_similarity(*a, **k):
    raise NotImplementedError


def peak_signal_noise_ratio(*a, **k):
    raise NotImplementedError
