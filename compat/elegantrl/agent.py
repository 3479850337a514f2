"""elegantrl/agent.py: the on-policy agent of the PIME path (AgentPPO :543-712).  The off-policy agents (DQN / DDPG / TD3 /
SAC) are outside the path this library accelerates; asking for one says so."""
from pime_b200.rl import AgentPPO  # noqa: F401


def _off_policy(name):
    class _Unavailable:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: off-policy agents are not part of the PIME hot path (DESIGN.md section 7); "
                                      "use ppo / residualppo / residualintegratormodularppo")
    _Unavailable.__name__ = name
    return _Unavailable


AgentTD3, AgentSAC, AgentDDPG, AgentDQN = (_off_policy(n) for n in ("AgentTD3", "AgentSAC", "AgentDDPG", "AgentDQN"))
