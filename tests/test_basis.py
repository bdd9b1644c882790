"""CPU tests for the product-side basis generator and the device-code generator."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN
from dig_b200 import basis, codegen

with open(os.path.join(GOLDEN, "basis_formulas.json")) as _fh:
    GOLD = json.load(_fh)


@pytest.mark.parametrize("tag,flavor,ns,nr", [("spherenet_7_6", "dimenet", 7, 6),
                                             ("spherenet_3_6", "dimenet", 3, 6),
                                             ("comenet_2_3", "gemnet", 2, 3),
                                             ("pronet_2_6", "gemnet", 2, 6)])
def test_generator_reproduces_reference_formulas(tag, flavor, ns, nr):
    """dig_b200/basis.py must emit exactly the strings the reference's sympy code lambdifies."""
    mine = basis.basis_sources(flavor, ns, nr)
    ref = GOLD[tag]
    assert mine["bessel"] == ref["bessel"]
    for key in ("yl0", "ylm"):
        # the l = 0 entry is a constant the reference stores through a float32 tensor
        assert np.float32(float(mine[key][0])) == np.float32(float(ref[key][0]))
        assert mine[key][1:] == ref[key][1:]


def test_dimenetpp_uses_the_same_forms_as_spherenet():
    assert GOLD["dimenetpp_7_6"]["bessel"] == GOLD["spherenet_7_6"]["bessel"]
    assert GOLD["dimenetpp_7_6"]["yl0"] == GOLD["spherenet_7_6"]["yl0"]


def test_committed_headers_match_generator():
    csrc = os.path.join(os.path.dirname(basis.__file__), "csrc", "generated")
    for tag, (flavor, ns, nr) in codegen.CONFIGS.items():
        want = codegen.emit_header(tag, flavor, ns, nr, basis.basis_sources(flavor, ns, nr))
        with open(os.path.join(csrc, f"basis_{tag}.cuh")) as fh:
            assert fh.read() == want, f"stale generated header for {tag}: run python -m dig_b200.codegen --force"


def test_codegen_typing_rules():
    src = codegen.emit_function("f", ["x"], ["2*x/5 + x**2 - 1/x + x**7 + (3/2)*x**0.5"])
    assert "__fmul_rn(2.0f, x)" in src                     # python scalar rounded to fp32, one op per node
    assert "0.200000003f" in src                           # t / c -> t * (1.0f / c)  (ATen CUDA)
    assert "__fmul_rn(x, x)" in src                        # ** 2 -> x*x
    assert "__fdiv_rn(1.0f, x)" in src                     # c / t -> reciprocal(t) * c
    assert "powf(x, 7.0f)" in src and "sqrtf(x)" in src
    assert "1.5f" in src                                   # (3/2) folded in python double first
    assert "fmaf" not in src                               # never contracted


def test_envelope_coefficients():
    assert basis.envelope_coefficients(5) == (6, -28.0, 48, -21.0)


@pytest.mark.parametrize("flavor,ns,nr", [("dimenet", 7, 6), ("dimenet", 3, 6)])
def test_derivative_sources_match_finite_differences(flavor, ns, nr):
    """The force path's generated derivatives (bessel_dx, yl0_dtheta, ylm_dtheta, ylm_dphi) are the symbolic
    derivatives of the reference's closed forms: check every one against a central difference of the VALUE source,
    evaluated in float64 (the value sources themselves are pinned to the reference above)."""
    import math
    src = basis.basis_sources(flavor, ns, nr)
    env = {"sin": math.sin, "cos": math.cos, "sqrt": math.sqrt, "pi": math.pi}

    def ev(s, **kw):
        return float(eval(s, dict(env, **kw)))

    h = 1e-6
    for x in (0.31, 0.77):
        for v, d in zip(src["bessel"], src["bessel_dx"]):
            fd = (ev(v, x=x + h) - ev(v, x=x - h)) / (2 * h)
            an = ev(d, x=x)
            assert abs(an - fd) <= 1e-5 * max(1.0, abs(fd)) + 1e-6 * max(abs(ev(v, x=x)) / h * 1e-9, 1.0), (v[:40], x)
    for th in (0.4, 1.9):
        for v, d in zip(src["yl0"], src["yl0_dtheta"]):
            fd = (ev(v, theta=th + h) - ev(v, theta=th - h)) / (2 * h)
            assert abs(ev(d, theta=th) - fd) <= 1e-6 * max(1.0, abs(fd))
        for ph in (0.3, 4.1):
            for v, dt, dp in zip(src["ylm"], src["ylm_dtheta"], src["ylm_dphi"]):
                fdt = (ev(v, theta=th + h, phi=ph) - ev(v, theta=th - h, phi=ph)) / (2 * h)
                fdp = (ev(v, theta=th, phi=ph + h) - ev(v, theta=th, phi=ph - h)) / (2 * h)
                assert abs(ev(dt, theta=th, phi=ph) - fdt) <= 1e-6 * max(1.0, abs(fdt))
                assert abs(ev(dp, theta=th, phi=ph) - fdp) <= 1e-6 * max(1.0, abs(fdp))
    assert len(src["bessel_dx"]) == ns * nr and len(src["ylm_dphi"]) == ns * ns


def test_second_order_sources_match_finite_differences():
    """basis_sources_second_order (groundwork for force training through the angular basis): second derivatives vs a
    central difference of the FIRST-derivative sources, in float64."""
    import math
    first = basis.basis_sources("dimenet", 3, 6)
    second = basis.basis_sources_second_order("dimenet", 3, 6)
    env = {"sin": math.sin, "cos": math.cos, "sqrt": math.sqrt, "pi": math.pi}

    def ev(s, **kw):
        return float(eval(s, dict(env, **kw)))

    h = 1e-6
    for x in (0.37, 0.81):
        for d1, d2 in zip(first["bessel_dx"], second["bessel_dxx"]):
            fd = (ev(d1, x=x + h) - ev(d1, x=x - h)) / (2 * h)
            assert abs(ev(d2, x=x) - fd) <= 2e-5 * max(1.0, abs(fd))
    for th in (0.5, 2.2):
        for d1, d2 in zip(first["yl0_dtheta"], second["yl0_dtheta2"]):
            fd = (ev(d1, theta=th + h) - ev(d1, theta=th - h)) / (2 * h)
            assert abs(ev(d2, theta=th) - fd) <= 1e-6 * max(1.0, abs(fd))
    assert len(second["bessel_dxx"]) == 18 and len(second["yl0_dtheta2"]) == 3


@pytest.mark.parametrize("ns", [7, 3])
def test_recurrence_harmonics_equal_the_closed_forms(ns):
    """csrc/harmonics.cuh (numpy twin: basis.harmonics_recurrence) evaluates the harmonics of the fused projection from
    the recurrences the reference's symbolic construction starts from; in exact arithmetic they ARE the reference's
    simplified closed forms (flat order included), and in fp32 they are at least as close to the fp64 values."""
    import math
    src = basis.basis_sources("dimenet", ns, 6)
    rng = np.random.default_rng(0)
    th, ph = rng.uniform(0, math.pi, 2000), rng.uniform(0, 2 * math.pi, 2000)
    env = {"sin": np.sin, "cos": np.cos, "sqrt": np.sqrt, "pi": math.pi}
    ref = np.stack([np.broadcast_to(eval("lambda theta, phi: " + s, env)(th, ph), th.shape) for s in src["ylm"]], -1)
    got = basis.harmonics_recurrence(th, ph, ns)
    assert got.shape == ref.shape == (2000, ns * ns)
    assert np.abs(got - ref).max() < 1e-12
    y0 = np.stack([np.broadcast_to(eval("lambda theta: " + s, env)(th), th.shape) for s in src["yl0"]], -1)
    assert np.abs(y0 - got[:, [l * l for l in range(ns)]]).max() < 1e-12       # the zonal entries double as yl0
    got32 = basis.harmonics_recurrence(th, ph, ns, np.float32)
    env32 = dict(env, pi=np.float32(math.pi))
    th32, ph32 = th.astype(np.float32), ph.astype(np.float32)
    ref32 = np.stack([np.broadcast_to(eval("lambda theta, phi: " + s, env32)(th32, ph32), th.shape)
                      for s in src["ylm"]], -1)
    assert np.abs(got32 - ref).max() < 3e-6
    assert np.abs(got32 - ref).max() <= 1.5 * np.abs(ref32.astype(np.float64) - ref).max() + 5e-7
