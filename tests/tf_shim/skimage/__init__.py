"""Import stub (tests/tf_shim): lets the reference modules import where this third-party package is absent. TEST INFRASTRUCTURE."""
