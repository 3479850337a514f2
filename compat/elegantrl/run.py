"""elegantrl/run.py:14-225,478-619 (single-process trainer, evaluator)."""
from pime_b200.rl import (Arguments, Evaluator, ReplayBuffer, PreprocessEnv, evaluate_batched, get_episode_return,  # noqa: F401
                          train_and_evaluate)
