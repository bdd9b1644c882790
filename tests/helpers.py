"""Shared test utilities (test infrastructure; may import oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle.gen_golden import CASES  # noqa: E402  (case table only; no reference import)
from oracle.weights import formula_state_dict  # noqa: E402


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def case_inputs(name, device="cpu"):
    g = load_golden(name)
    z = torch.from_numpy(g["z"]).to(device)
    pos = torch.from_numpy(g["pos"]).to(device)
    batch = torch.from_numpy(g["batch"]).to(device)
    return g, z, pos, batch


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
