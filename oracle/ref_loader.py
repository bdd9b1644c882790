"""TEST INFRASTRUCTURE ONLY.  Import the reference's own model code, verbatim.

Works only where /root/reference exists (the build container).  On the GPU box the
golden fixtures under tests/golden/ and oracle/restated.py stand in for it.
"""
import math
import os
import sys

REFERENCE_ROOT = os.environ.get("DIG_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dig", "threedgraph", "method"))


def load_reference():
    """Returns the module `dig.threedgraph.method` of the reference (SchNet, SphereNet,
    DimeNetPP, ComENet, run) running over oracle/shim.py."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import shim
    shim.install()
    import numpy as np
    if not hasattr(np, "math"):  # numpy>=2 dropped np.math; used at spherenet/features.py:70
        np.math = math
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dig.threedgraph.method as method  # noqa: E402
    return method


def load_reference_utils():
    load_reference()
    import dig.threedgraph.utils as utils
    return utils
