"""Import the unmodified reference package from oracle/_ref (see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY.

`load()` returns the `ssp_bayes_opt` module, or None when the copy is absent.  The package imports `nengo` unconditionally
(ssp_bayes_opt/__init__.py:15-16) and nengo is not installed: the module is stubbed in sys.modules first;
the stub is never called on the hot path (sspspace.py, blr.py, agents/ssp_agent.py use numpy / scipy / scikit-learn only).
"""
import os
import sys
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "ssp_bayes_opt", "sspspace.py"))


def load():
    if not available():
        return None
    for name in ("nengo", "nengo.params", "nengo.utils", "nengo.utils.builder", "nengo.dists",
                 "nengo.processes", "nengo.solvers", "nengo.neurons"):
        sys.modules.setdefault(name, MagicMock())
    sys.modules["nengo"].Network = type("Network", (object,), {})
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ssp_bayes_opt  # noqa: F401
    return ssp_bayes_opt
