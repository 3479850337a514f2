"""elegantrl/agent_residual.py:15-98."""
from pime_b200.rl import AgentResidualIntegratorModularPPO, AgentResidualPPO, Residual  # noqa: F401
