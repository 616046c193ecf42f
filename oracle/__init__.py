"""CPU oracle for the SSP-BO hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this package.  The product (`ssp_bayesopt_b200`) never does.
"""
