"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def host(t):
    return t.detach().cpu().numpy()


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def random_wt_inputs(rng, n):
    d = dict(a1=rng.uniform(0.0015, 0.0024, n), a2=rng.uniform(0.0015, 0.0024, n), Kp=rng.uniform(0.07, 0.17, n),
             h1=rng.uniform(0, 12, n), h2=rng.uniform(0, 12, n), r=rng.uniform(0, 10, n), I=rng.uniform(-25, 25, n),
             t=rng.integers(0, 199, n).astype(np.int32), action=rng.uniform(-1.5, 1.5, n),
             noise1=rng.normal(0, 0.01, n), noise2=rng.normal(0, 0.01, n))
    d["h1"][: n // 16] = 0.0
    d["h2"][n // 32: n // 8] = 0.0
    d["t"][-4:] = 199
    return d


def load_wt(env, d):
    """Copy a dict of numpy arrays into a WaterTankVec's state tensors."""
    for k in ("h1", "h2", "r", "I", "a1", "a2", "Kp"):
        getattr(env, k).copy_(dev(d[k], env.dtype))
    env.t.copy_(dev(d["t"], torch.int32))
    env.ep_return.zero_()


def load_ph(env, d):
    for k in ("x", "y", "r", "I", "A", "B", "C", "qww_V", "qc_V"):
        if k in d:
            getattr(env, k).copy_(dev(d[k], getattr(env, k).dtype))   # x, A, B are fp64 in both flavours
    env.t.copy_(dev(d["t"], torch.int32))
    env.ep_return.zero_()
