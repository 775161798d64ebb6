"""GPU parity of the chain against the oracle and the golden fixture, through the C ABI."""
import numpy as np
import pytest

import synth
from helpers import assert_parity, compare_layer

pytestmark = pytest.mark.gpu


def _both(te, ob, rows, cols, res=0.02, position=(0.0, 0.0)):
    return te.Geometry.make(rows, cols, res, position), ob.Geometry.make(rows, cols, res, position)


@pytest.mark.parametrize("kernel", ["generic", "auto"])
def test_fixture_through_gpu(te, ctx, oracle, fixture_map, kernel):
    m, d = fixture_map
    g = te.Geometry(m["rows"], m["cols"], m["resolution"], m["length_x"], m["length_y"], *m["position"], 0, 0)
    ctx.set_kernel(te.KERNEL_GENERIC if kernel == "generic" else te.KERNEL_AUTO)
    o = ctx.chain_host(g, te.ChainParams.yaml_defaults(0), d["elevation"])
    ref = {"slope": d["traversability_slope"], "step": d["traversability_step"],
           "roughness": d["traversability_roughness"], "traversability": d["traversability"]}
    reports = assert_parity(o, ref)
    if kernel == "generic":
        for r in reports:  # the literal kernel replays the reference's double arithmetic
            assert r["bit_exact"] >= r["cells"] - 2, r


@pytest.mark.parametrize("kernel", ["generic", "auto"])
@pytest.mark.parametrize("case", [
    dict(rows=96, cols=80, seed=1, preset="gentle"),
    dict(rows=200, cols=168, seed=2, preset="mixed"),
    dict(rows=160, cols=130, seed=3, preset="rough"),
    dict(rows=150, cols=140, seed=4, preset="mixed", position=(123.456, -78.9)),
    dict(rows=120, cols=133, seed=5, preset="mixed", res=0.03),
])
def test_chain_matches_oracle(te, ctx, oracle, case, kernel):
    res = case.get("res", 0.02)
    pos = case.get("position", (0.0, 0.0))
    z = synth.terrain(case["rows"], case["cols"], res, case["seed"], case["preset"], pos)
    g, og = _both(te, oracle, case["rows"], case["cols"], res, pos)
    ref = oracle.chain(og, oracle.ChainParams.yaml_defaults(0), z, with_normals=True)
    ctx.set_kernel(te.KERNEL_GENERIC if kernel == "generic" else te.KERNEL_AUTO)
    got = ctx.chain_host(g, te.ChainParams.yaml_defaults(0), z, with_normals=True)
    reports = assert_parity(got, ref)
    for k in ("nx", "ny", "nz"):
        r = compare_layer(got[k], ref[k], k)
        assert r["nan_mismatch"] == 0 and r["max_abs"] < 1e-6, r
    assert any(np.isnan(ref["slope"]).ravel()) == any(np.isnan(z).ravel())
    print(reports)
