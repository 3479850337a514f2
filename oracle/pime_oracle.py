"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/pime_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module.  The product package never does (tests/test_product_isolation.py greps for it).

Parity status: PINNED against outputs of the reference itself (tests/golden/, produced by
oracle/gen_golden.py) -- see the header of pime_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpime_oracle.so")

REWARD = {"distance": 0, "square_distance": 1, "sparse": 2}


class WtCfg(C.Structure):
    _fields_ = [("A1", C.c_double), ("A2", C.c_double), ("G", C.c_double), ("sample_t", C.c_double),
                ("n_discrete", C.c_int32), ("max_step", C.c_int32), ("P_max_action", C.c_double),
                ("reward_type", C.c_int32), ("has_integrator", C.c_int32), ("z1", C.c_double),
                ("distance_threshold", C.c_double), ("integral_max", C.c_double), ("integral_punish", C.c_double)]


class PhCfg(C.Structure):
    _fields_ = [("reward_type", C.c_int32), ("integrator_mode", C.c_int32), ("max_episode_steps", C.c_int32),
                ("table_len", C.c_int32), ("act_low", C.c_double), ("act_high", C.c_double),
                ("distance_threshold", C.c_double), ("integral_max", C.c_double), ("integral_punish", C.c_double),
                ("action_punishment", C.c_double), ("action_change_punishment", C.c_double), ("mhcl_step", C.c_double)]


class ActorCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("state_dim", C.c_int32), ("mid_dim", C.c_int32), ("integrator_dim", C.c_int32)]


def wt_cfg(reward_type="square_distance", has_integrator=True, **kw) -> WtCfg:
    """Registered water-tank config (reference gym_control/__init__.py:29-69)."""
    d = dict(A1=1.0, A2=1.0, G=980.0, sample_t=2.0, n_discrete=20, max_step=200, P_max_action=10.0,
             reward_type=REWARD[reward_type], has_integrator=int(has_integrator), z1=1.0, distance_threshold=0.05,
             integral_max=25.0, integral_punish=0.0)
    d.update(kw)
    return WtCfg(**d)


def ph_cfg(reward_type="square_distance", integrator_mode=1, **kw) -> PhCfg:
    """Registered pH config (reference gym_control/__init__.py:3-27, ph.py:26-50,146-147)."""
    d = dict(reward_type=REWARD[reward_type], integrator_mode=integrator_mode, max_episode_steps=50, table_len=100000,
             act_low=0.0, act_high=1.5, distance_threshold=0.05, integral_max=25.0, integral_punish=0.0,
             action_punishment=0.0, action_change_punishment=0.0, mhcl_step=1e-5)
    d.update(kw)
    return PhCfg(**d)


_lib = None
_lock = threading.Lock()


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pime_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    with _lock:
        if _lib is None:
            _lib = C.CDLL(build())
            _lib.pime_oracle_ph_index.restype = C.c_int32
            _lib.pime_oracle_ph_step.restype = C.c_int64
            _lib.pime_oracle_ph_rollout.restype = C.c_int64
            _lib.pime_oracle_actor_param_count.restype = C.c_int64
    return _lib


def _p(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (type(a), getattr(a, "dtype", None))
    return a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def wt_step(cfg, h1, h2, r, I, t, a1, a2, Kp, action, noise1=None, noise2=None):
    """In place on h1,h2,I,t (float64/int32 arrays). Returns (reward, done)."""
    n = h1.shape[0]
    reward = np.empty(n, np.float64)
    done = np.empty(n, np.uint8)
    lib().pime_oracle_wt_step(C.byref(cfg), C.c_int64(n), _p(h1, np.float64), _p(h2, np.float64), _p(r, np.float64),
                              _p(I, np.float64), _p(t, np.int32), _p(a1, np.float64), _p(a2, np.float64),
                              _p(Kp, np.float64), _p(action, np.float64), _p(noise1, np.float64), _p(noise2, np.float64),
                              _p(reward, np.float64), _p(done, np.uint8))
    return reward, done


CHEM = dict(kw=1e-14, kchem=5.6e-10, ka=0.5e-5, MNaOH=0.01, MHA=0.005, MNH3=0.01)


def ph_table(table_len=100000, mhcl_step=1e-5, npow=0, **chem):
    c = dict(CHEM)
    c.update(chem)
    out = np.empty(table_len, np.float64)
    lib().pime_oracle_ph_table(C.c_int32(table_len), C.c_double(mhcl_step), C.c_double(c["kw"]), C.c_double(c["kchem"]),
                               C.c_double(c["ka"]), C.c_double(c["MNaOH"]), C.c_double(c["MHA"]), C.c_double(c["MNH3"]),
                               C.c_int32(npow), _p(out, np.float64))
    return out


def ph_update_system(qww_V, qc_V, sample_t=20.0):
    qww_V, qc_V = f64(qww_V), f64(qc_V)
    n = qww_V.shape[0]
    A, B, Cc = (np.empty(n, np.float64) for _ in range(3))
    lib().pime_oracle_ph_update_system(C.c_int64(n), C.c_double(sample_t), _p(qww_V, np.float64), _p(qc_V, np.float64),
                                       _p(A, np.float64), _p(B, np.float64), _p(Cc, np.float64))
    return A, B, Cc


def ph_index(cfg, Cx: float) -> int:
    return int(lib().pime_oracle_ph_index(C.byref(cfg), C.c_double(Cx)))


def ph_step(cfg, table, x, y, r, I, t, A, B, Cc, action):
    n = x.shape[0]
    reward = np.empty(n, np.float64)
    done = np.empty(n, np.uint8)
    err = lib().pime_oracle_ph_step(C.byref(cfg), _p(table, np.float64), C.c_int64(n), _p(x, np.float64),
                                    _p(y, np.float64), _p(r, np.float64), _p(I, np.float64), _p(t, np.int32),
                                    _p(A, np.float64), _p(B, np.float64), _p(Cc, np.float64), _p(action, np.float64),
                                    _p(reward, np.float64), _p(done, np.uint8))
    if err:
        raise IndexError(f"pH table lookup out of range for env {err - 1} (reference ph.py:188 raises IndexError)")
    return reward, done


# ---------------------------------------------------------------------------------------------- actor
MODULAR_KEYS = ["other_net.0.weight", "other_net.0.bias", "other_net.2.weight", "other_net.2.bias",
                "integrator_net.0.weight", "integrator_net.0.bias", "integrator_net.2.weight", "integrator_net.2.bias",
                "net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias"]
PLAIN_KEYS = ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "net.4.weight", "net.4.bias",
              "net.6.weight", "net.6.bias"]


def pack_actor_params(state_dict, kind: int) -> np.ndarray:
    """Flatten a reference actor state_dict (torch tensors or arrays) into the oracle's fp32 parameter pack."""
    keys = MODULAR_KEYS if kind == 1 else PLAIN_KEYS
    parts = []
    for k in keys:
        v = state_dict[k]
        v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        parts.append(np.ascontiguousarray(v, np.float32).reshape(-1))
    return np.concatenate(parts)


def actor_forward(acfg: ActorCfg, params: np.ndarray, obs: np.ndarray) -> np.ndarray:
    obs = np.ascontiguousarray(obs, np.float32)
    assert params.shape[0] == lib().pime_oracle_actor_param_count(C.byref(acfg)), "parameter pack size mismatch"
    out = np.empty(obs.shape[0], np.float32)
    lib().pime_oracle_actor_forward(C.byref(acfg), _p(params, np.float32), C.c_int64(obs.shape[0]), _p(obs, np.float32),
                                    _p(out, np.float32))
    return out


def wt_rollout(cfg, acfg, params, a_std_log, priorK, obs_mode, num_stack, deterministic, T, h1, h2, r, I, t, a1, a2, Kp,
               frames=None, eps=None, pn1=None, pn2=None, reward_scale=1.0, gamma=0.99, want_replay=True,
               want_actions=False):
    """Runs T steps in place; returns dict(buf_state, buf_other, ep_return, env_action)."""
    n = h1.shape[0]
    S = 3 if obs_mode == 0 else (4 if obs_mode == 1 else 3 * num_stack)
    buf_state = np.zeros((T, n, S), np.float32) if want_replay else None
    buf_other = np.zeros((T, n, 4), np.float32) if want_replay else None
    ep_return = np.zeros(n, np.float64)
    env_action = np.zeros((T, n), np.float64) if want_actions else None
    priorK = f64(priorK).reshape(-1)
    assert priorK.shape[0] == S
    lib().pime_oracle_wt_rollout(C.byref(cfg), C.byref(acfg) if acfg is not None else None, _p(params, np.float32),
                                 C.c_float(a_std_log), _p(priorK, np.float64), C.c_int32(obs_mode), C.c_int32(num_stack),
                                 C.c_int32(int(deterministic)), C.c_int64(n), C.c_int32(T), _p(h1, np.float64),
                                 _p(h2, np.float64), _p(r, np.float64), _p(I, np.float64), _p(t, np.int32),
                                 _p(a1, np.float64), _p(a2, np.float64), _p(Kp, np.float64), _p(frames, np.float64),
                                 _p(eps, np.float32), _p(pn1, np.float64), _p(pn2, np.float64), C.c_double(reward_scale),
                                 C.c_double(gamma), _p(buf_state, np.float32), _p(buf_other, np.float32),
                                 _p(ep_return, np.float64), _p(env_action, np.float64))
    return dict(buf_state=buf_state, buf_other=buf_other, ep_return=ep_return, env_action=env_action)


def ph_rollout(cfg, table, acfg, params, a_std_log, priorK, deterministic, T, x, y, r, I, t, A, B, Cc, eps=None,
               reward_scale=1.0, gamma=0.99, want_replay=True, want_actions=False):
    n = x.shape[0]
    S = 3 if cfg.integrator_mode else 2
    buf_state = np.zeros((T, n, S), np.float32) if want_replay else None
    buf_other = np.zeros((T, n, 4), np.float32) if want_replay else None
    ep_return = np.zeros(n, np.float64)
    env_action = np.zeros((T, n), np.float64) if want_actions else None
    priorK = f64(priorK).reshape(-1)
    err = lib().pime_oracle_ph_rollout(C.byref(cfg), _p(table, np.float64), C.byref(acfg) if acfg is not None else None,
                                       _p(params, np.float32), C.c_float(a_std_log), _p(priorK, np.float64),
                                       C.c_int32(int(deterministic)), C.c_int64(n), C.c_int32(T), _p(x, np.float64),
                                       _p(y, np.float64), _p(r, np.float64), _p(I, np.float64), _p(t, np.int32),
                                       _p(A, np.float64), _p(B, np.float64), _p(Cc, np.float64), _p(eps, np.float32),
                                       C.c_double(reward_scale), C.c_double(gamma), _p(buf_state, np.float32),
                                       _p(buf_other, np.float32), _p(ep_return, np.float64), _p(env_action, np.float64))
    if err:
        raise IndexError(f"pH table lookup out of range for env {err - 1}")
    return dict(buf_state=buf_state, buf_other=buf_other, ep_return=ep_return, env_action=env_action)


# ---------------------------------------------------------------------------------------------- RNG
def philox4x32(seed: int, index: int, tick: int, stream: int) -> np.ndarray:
    out = (C.c_uint32 * 4)()
    lib().pime_oracle_philox4x32(C.c_uint64(seed), C.c_uint64(index), C.c_uint32(tick), C.c_uint32(stream), out)
    return np.array(list(out), dtype=np.uint32)


def reset_uniforms(seed: int, index: int, episode: int) -> np.ndarray:
    out = (C.c_double * 6)()
    lib().pime_oracle_reset_uniforms(C.c_uint64(seed), C.c_uint64(index), C.c_uint32(episode), out)
    return np.array(list(out), dtype=np.float64)
