"""The C++ plugin shells (filters::FilterBase<grid_map::GridMap>) driving the GPU, checked against the golden fixture."""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "traversability_estimation_b200", "plugin")


def _load(prefix, name, rows, cols):
    return np.fromfile(f"{prefix}_{name}.bin", dtype=np.float32).reshape((cols, rows)).T


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("start", [None, (37, 50)])
def test_yaml_chain_through_plugins_on_fixture(fixture_map, tmp_path, start, fuse):
    m, d = fixture_map
    subprocess.check_call(["make", "-C", PLUGIN, "-s"])
    rows, cols = m["rows"], m["cols"]
    src = tmp_path / "elev.bin"
    np.ascontiguousarray(d["elevation"].T).tofile(src)
    out = str(tmp_path / "out")
    cmd = [os.path.join(PLUGIN, "test_plugins"), "chain", str(rows), str(cols), repr(m["resolution"]), "0", "0", str(src), out]
    if start:
        cmd += [str(start[0]), str(start[1])]
    env = dict(os.environ, TE_B200_FUSE_CHAIN="1" if fuse else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout
    ref = {"slope": d["traversability_slope"], "step": d["traversability_step"], "roughness": d["traversability_roughness"],
           "traversability": d["traversability"]}
    fused = {k: _load(out, f"fused_{k}", rows, cols) for k in ref}
    assert_parity(fused, ref)
    if not start:
        single = {k: _load(out, k, rows, cols) for k in ("slope", "step", "roughness")}
        if fuse:
            # UNCHANGED YAML: the first of the three reference-named plugins launched the fused chain once, all three layers
            # came from its cache, after the upstream surface normals were verified against ours
            assert "REGISTRY launches=1 served=3" in r.stdout, r.stdout
            assert_parity(single, ref, keys=("slope", "step", "roughness"))
        else:
            assert "REGISTRY launches=0 served=0" in r.stdout, r.stdout
            for k in single:  # the stand-alone filters run the literal kernels: bit-exact on the fixture
                assert np.array_equal(single[k].view(np.uint32), ref[k].view(np.uint32)), k
