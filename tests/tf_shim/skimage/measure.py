"""Import stub (tests/tf_shim): lets the reference modules import where this third-party package is absent. TEST INFRASTRUCTURE."""


def compare_ssim(*a, **k):
    raise NotImplementedError


def compare_psnr(*a, **k):
    raise NotImplementedError
