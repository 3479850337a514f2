"""elegantrl/utils.py:10-47."""
from pime_b200.rl import configure_logger, get_latest_run_id  # noqa: F401
