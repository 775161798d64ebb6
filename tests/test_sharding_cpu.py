"""Host-side logic of the multi-GPU column-slab tiling, exercised with world_size 2 and 3 over gloo on CPU."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_covers_map_without_overlap():
    from traversability_estimation_b200.sharding import plan_slab
    for cols, world, halo in ((8192, 8, 4), (133, 3, 4), (200, 4, 41), (64, 1, 4)):
        plans = [plan_slab(cols, world, r, halo) for r in range(world)]
        assert plans[0].col_begin == 0 and plans[0].halo_left == 0 and plans[-1].halo_right == 0
        assert sum(p.col_count for p in plans) == cols
        for a, b in zip(plans, plans[1:]):
            assert a.col_begin + a.col_count == b.col_begin
        for p in plans:
            assert p.halo_left == min(halo, p.col_begin)
            assert p.halo_right == min(halo, cols - p.col_begin - p.col_count)


def test_plan_rejects_slabs_narrower_than_the_halo():
    """The halo exchange is one hop: a rank that owns fewer columns than the halo cannot serve its neighbour."""
    from traversability_estimation_b200.sharding import plan_slab
    with pytest.raises(ValueError, match="narrower than the halo"):
        plan_slab(10, 4, 1, 41)
    with pytest.raises(ValueError, match="narrower than the halo"):
        plan_slab(8192, 8, 0, 1025)
    plan_slab(10, 1, 0, 41)  # a single rank exchanges nothing


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from traversability_estimation_b200.sharding import plan_slab, exchange_halo
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, cols, halo = 24, 37, 4
    rng = np.random.default_rng(5)
    full = rng.standard_normal((cols, rows)).astype(np.float32)   # (cols, rows): column-major layer
    full[7, 3] = np.nan
    p = plan_slab(cols, world, rank, halo)
    buf = torch.full((p.buffer_cols, rows), float("nan"))
    buf[p.halo_left:p.halo_left + p.col_count] = torch.from_numpy(full[p.col_begin:p.col_begin + p.col_count])
    exchange_halo(dist, buf, p, halo)
    want = full[p.col_begin - p.halo_left:p.col_begin + p.col_count + p.halo_right]
    ok = np.array_equal(buf.numpy(), want, equal_nan=True)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)
""")


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
