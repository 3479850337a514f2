"""elegantrl/replay.py:238-379 (the on-policy ReplayBuffer, storage in HBM)."""
from pime_b200.rl import ReplayBuffer  # noqa: F401
