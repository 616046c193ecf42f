from .agent import Agent
from .ssp_agent import SSPAgent
from .ssp_traj_agent import SSPTrajectoryAgent
from .ssp_multi_agent import SSPMultiAgent
from .rff_agent import RFFAgent
from . import domains

__all__ = ["Agent", "SSPAgent", "SSPTrajectoryAgent", "SSPMultiAgent", "RFFAgent", "domains"]
