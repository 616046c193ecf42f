"""ssp_bayesopt_b200: the data-parallel hot path of SSP-BO (ctn-waterloo/ssp-bayesopt) on NVIDIA B200.

Batched SSP encoding of candidate sets fused with BLR-UCB/MI acquisition scoring and argmax, from-set
decoding and the float64 rank-one posterior update, as hand-written sm_100a CUDA behind a C ABI
(include/sspbo.h), re-hosted under the reference's Python names.  See DESIGN.md.
"""
from . import _lib, agents, blr, dist, sspspace
from .bayesian_optimization import BayesianOptimization
from .blr import BayesianLinearRegression
from .engine import DeviceEngine
from .sspspace import HexagonalSSPSpace, RandomSSPSpace, SPSpace, SSPSpace
from .agents import RFFAgent, SSPAgent, SSPMultiAgent, SSPTrajectoryAgent
from .batched import BatchedRuns

__all__ = ["BayesianOptimization", "BayesianLinearRegression", "DeviceEngine", "SSPSpace", "HexagonalSSPSpace",
           "RandomSSPSpace", "SPSpace", "SSPAgent", "SSPTrajectoryAgent", "SSPMultiAgent", "RFFAgent", "BatchedRuns", "agents", "blr", "dist", "sspspace"]
