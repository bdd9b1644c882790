"""CPU tests: the C-ABI library loads and exports every symbol include/dig3d.h declares; the
product path fails loudly without CUDA (no fallback)."""
import os
import re

import pytest
import torch

from helpers import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "dig3d.h")) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"\b(dig3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dig_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libdig3d.so lacks {name}"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.dig3d_abi_version() == 1


def test_library_is_sm100a():
    import subprocess
    from dig_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    from dig_b200.threedgraph.method import DimeNetPP
    from dig_b200.data import synthetic_batch
    model = DimeNetPP()
    batch = synthetic_batch(2, "qm9", seed=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(batch)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dig_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                assert "from .. import oracle" not in src


def test_state_dict_contract():
    """Parameter names / shapes equal the reference's, from the fixture."""
    import json
    from helpers import GOLDEN
    from dig_b200.threedgraph.method import SphereNet, DimeNetPP, SchNet, ComENet
    with open(os.path.join(GOLDEN, "state_shapes.json")) as fh:
        shapes = json.load(fh)
    for cls, kw in ((SphereNet, dict(cutoff=5.0)), (DimeNetPP, dict(cutoff=5.0)),
                    (SphereNet, dict(cutoff=5.0, num_spherical=3)),
                    (SchNet, dict(num_layers=2, hidden_channels=32, num_filters=32, cutoff=10.0)),
                    (ComENet, dict(cutoff=6.0, hidden_channels=256, middle_channels=64))):
        ref = shapes[cls.__name__ + json.dumps(kw, sort_keys=True)]
        mine = {k: list(v.shape) for k, v in cls(**kw).state_dict().items()}
        assert mine == ref


def test_evaluator_known_answer():
    # reference test/threedgraph/evaluation/test_ThreeDEvaluator.py:7-21
    import numpy as np
    from dig_b200.threedgraph.evaluation import ThreeDEvaluator
    ev = ThreeDEvaluator()
    assert ev.eval({"y_true": np.array([1.0, -0.5]), "y_pred": np.array([0.6, 0.0])})["mae"] == pytest.approx(0.45)
    assert ev.eval({"y_true": torch.tensor([1.0, -0.5]), "y_pred": torch.tensor([0.6, 0.0])})["mae"] == pytest.approx(0.45)
