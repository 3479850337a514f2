"""The reference's ``elegantrl`` package name on top of pime_b200.rl (see compat/README.md)."""
from pime_b200 import logger  # noqa: F401  (`from elegantrl import logger`: agent.py:11, run.py:12, utils.py:2, train.py:17)
