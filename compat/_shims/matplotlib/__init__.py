"""No-op stand-in for matplotlib (absent from this image): train.py and utils/test.py import pyplot at module level."""
__version__ = "0+pime-shim"


def use(*a, **k):
    pass
