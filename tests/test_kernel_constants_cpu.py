"""Host-side checks of numeric constants that the CUDA kernels hard-code (no GPU needed)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUSED = os.path.join(ROOT, "traversability_estimation_b200", "csrc", "te_fused.cu")


def _acos_coefficients():
    src = open(FUSED).read()
    got = dict(re.findall(r"a\.k_p([0-7]) = B2\(([-0-9.eE]+)\)", src))
    assert sorted(got) == [str(k) for k in range(8)], got
    return [np.float32(float(got[str(k)])) for k in range(8)]


def test_fused_acos_polynomial_meets_its_stated_error():
    """acos2 in te_fused.cu: acos(x) = sqrt(1 - x) * P7(x) on [0, 1], evaluated in float32 (Horner, as the kernel does).
    The slope layer is 1 - acos(n_z)/crit with crit = 1 rad, tolerance 1e-5 relative + 1e-6: the polynomial has to stay far
    inside that, and te_fused.cu documents 2.2e-7 absolute / 1.7e-7 relative."""
    p = _acos_coefficients()
    x = np.concatenate([np.linspace(0.0, 1.0, 2_000_001), 1.0 - np.logspace(-7, -1, 200_001)]).astype(np.float32)
    acc = np.full_like(x, p[7])
    for k in range(6, -1, -1):
        acc = (acc * x + p[k]).astype(np.float32)  # FFMA rounds once, mul+add twice: this is the pessimistic evaluation
    approx = (np.sqrt((np.float32(1.0) - x).astype(np.float32)).astype(np.float32) * acc).astype(np.float32)
    exact = np.arccos(x.astype(np.float64))
    err = np.abs(approx.astype(np.float64) - exact)
    assert err.max() < 4.0e-7, err.max()
    rel = err[exact > 1e-3] / exact[exact > 1e-3]
    assert rel.max() < 3.0e-7, rel.max()
    assert approx[x == 1.0].max() == 0.0  # flat cells: slope layer exactly 1


def test_footprint_slope_threshold_constant():
    """checkForSlope's nSlopesCritical = floor(2 * (3 res) * (maxGapWidth / 3) / res^2) (TraversabilityMap.cpp:871-873) is
    29 at 0.02 m — the double expression evaluates to 29.999999999999996 — and 20 at 0.03 m; the oracle and the CUDA
    predicate kernel must evaluate it in this operand order."""
    def ncrit(res, gap=0.3):
        return int(np.floor(2.0 * (3.0 * res) * (gap / 3.0) / (res * res)))
    assert ncrit(0.02) == 29 and ncrit(0.03) == 20
