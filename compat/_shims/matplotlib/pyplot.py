_PIME_SHIM = True


def _noop(*a, **k):
    return None


def __getattr__(name):   # plt.clf / plot / legend / savefig / close / figure ... : accepted and ignored
    return _noop
