"""utils/utils.py: name -> agent tables.  td3 / sac resolve to classes that explain they are out of scope."""
from elegantrl.agent import AgentPPO, AgentSAC, AgentTD3
from elegantrl.agent_residual import AgentResidualIntegratorModularPPO, AgentResidualPPO

MODELS = {"td3": AgentTD3, "ppo": AgentPPO, "sac": AgentSAC,
          "residualintegratormodularppo": AgentResidualIntegratorModularPPO, "residualppo": AgentResidualPPO}
IF_ONPOLICY = {"td3": False, "ppo": True, "sac": False, "residualintegratormodularppo": True, "residualppo": True}
