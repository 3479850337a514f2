"""elegantrl/env.py:10-72,194-245."""
from pime_b200.rl import PreprocessEnv  # noqa: F401


def get_gym_env_info(env, if_print=True):
    """(env_name, state_dim, action_dim, action_max, max_step, if_discrete, target_return) -- env.py:194-245."""
    p = PreprocessEnv(env, if_print=if_print)
    return p.env_name, p.state_dim, p.action_dim, p.action_max, p.max_step, p.if_discrete, p.target_return
