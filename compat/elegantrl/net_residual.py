"""elegantrl/net_residual.py: ActorResidualPPO (:6-66), ActorResidualIntegratorModularPPO (:138-205)."""
from pime_b200.rl import ActorResidualIntegratorModularPPO, ActorResidualPPO  # noqa: F401
