"""``elegantrl/logger.py`` is missing from the reference checkout; this is the module its call sites expect."""
from pime_b200.logger import *  # noqa: F401,F403
from pime_b200.logger import (Figure, close, configure, debug, dump, error, get_dir, history, info, log, record,  # noqa: F401
                              record_mean, set_level, values, warn)
