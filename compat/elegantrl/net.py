"""elegantrl/net.py: ActorPPO (:113-172), CriticAdv (:256-302), layer_norm (:617-619)."""
from pime_b200.rl import ActorPPO, CriticAdv, layer_norm  # noqa: F401
