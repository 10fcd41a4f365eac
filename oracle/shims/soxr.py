"""Import shim: the reference does `import soxr` at module top (inference.py:4) but only
calls it when sr != 22050 (inference.py:274-275).  libsoxr is not installed here.
TEST INFRASTRUCTURE ONLY."""


def resample(signal, in_rate, out_rate):
    raise RuntimeError("soxr is not available in this container (oracle shim)")
