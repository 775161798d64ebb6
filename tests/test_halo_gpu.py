"""te_halo_pull: the peer-mapped halo exchange behind the C ABI.  One GPU is enough: the "ranks" are slab
buffers of one process, or two processes that map each other's buffers through CUDA IPC on the same device."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_halo_pull_makes_slabs_bit_identical_to_the_whole_map(te, ctx):
    import torch
    rows, cols, H = 192, 250, 4
    z = synth.terrain(rows, cols, 0.02, 31, "mixed")
    g, p = te.Geometry.make(rows, cols, 0.02), te.ChainParams.yaml_defaults(0)
    ctx.set_kernel(te.KERNEL_AUTO)
    ctx.set_stream(None)
    zd = torch.from_numpy(np.ascontiguousarray(z.T)).cuda()
    whole = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g, p, zd, *whole, te.MEM_DEVICE)
    cuts = (0, 90, 160, 250)
    slabs, bufs = [], []
    for b, e in zip(cuts, cuts[1:]):
        hl, hr = min(H, b), min(H, cols - e)
        buf = torch.full((hl + (e - b) + hr, rows), float("nan"), dtype=torch.float32, device="cuda")
        buf[hl:hl + e - b] = zd[b:e]                      # only the owned columns: the halos come from te_halo_pull
        slabs.append(te.Slab(b, e - b, hl, hr))
        bufs.append(buf)
    torch.cuda.synchronize()
    ready, _ = ctx.event_create_ipc()
    ctx.event_record(ready)
    for k, (s, buf) in enumerate(zip(slabs, bufs)):
        left = te.HaloPeer(bufs[k - 1].data_ptr(), slabs[k - 1], ready) if k > 0 else None
        right = te.HaloPeer(bufs[k + 1].data_ptr(), slabs[k + 1], ready) if k + 1 < len(slabs) else None
        ctx.halo_pull(g, s, buf, left, right)
        outs = [torch.empty((s.col_count, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
        ctx.chain(g, p, buf, *outs, te.MEM_DEVICE, slab=s)
        ctx.synchronize()
        assert torch.equal(buf.view(torch.int32), zd[s.col_begin - s.halo_left:s.col_begin + s.col_count + s.halo_right].view(torch.int32))
        for w, o in zip(whole, outs):
            assert torch.equal(w[s.col_begin:s.col_begin + s.col_count].view(torch.int32), o.view(torch.int32)), k
    # a neighbour that owns fewer columns than the halo needs is refused (one-hop exchange), so is a missing neighbour
    with pytest.raises(te.TEError):
        ctx.halo_pull(g, te.Slab(100, 50, 4, 0), bufs[1], te.HaloPeer(bufs[0].data_ptr(), te.Slab(98, 2, 0, 0), None), None)
    with pytest.raises(te.TEError):
        ctx.halo_pull(g, slabs[1], bufs[1], None, None)
    ctx.event_destroy(ready)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tools"))
    import synth
    import traversability_estimation_b200 as te
    from traversability_estimation_b200.sharding import PeerHalo, plan_slab
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)                      # both processes share the one GPU: CUDA IPC maps across processes
    rows, cols, H = 128, 200, 4
    z = torch.from_numpy(np.ascontiguousarray(synth.terrain(rows, cols, 0.02, 77, "mixed").T)).cuda()
    g, p = te.Geometry.make(rows, cols, 0.02), te.ChainParams.yaml_defaults(0)
    ctx = te.Context(0)
    plan = plan_slab(cols, world, rank, H)
    buf = torch.full((plan.buffer_cols, rows), float("nan"), dtype=torch.float32, device="cuda")
    buf[plan.halo_left:plan.halo_left + plan.col_count] = z[plan.col_begin:plan.col_begin + plan.col_count]
    torch.cuda.synchronize()
    ph = PeerHalo(dist, ctx, te, buf, plan)
    ph.publish(); ctx.synchronize(); dist.barrier()
    ph.pull(g)
    outs = [torch.empty((plan.col_count, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g, p, buf, *outs, te.MEM_DEVICE, slab=ph.slab)
    ctx.synchronize()
    whole = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g, p, z, *whole, te.MEM_DEVICE)
    ctx.synchronize()
    ok = all(torch.equal(w[plan.col_begin:plan.col_begin + plan.col_count].view(torch.int32), o.view(torch.int32)) for w, o in zip(whole, outs))
    ok = ok and torch.equal(buf.view(torch.int32), z[plan.col_begin - plan.halo_left:plan.col_begin + plan.col_count + plan.halo_right].view(torch.int32))
    dist.barrier(); ph.close(); ctx.close()
    print("HALO_OK" if ok else "HALO_BAD", rank, flush=True)
    dist.destroy_process_group()
""")


def test_two_processes_pull_halos_through_cuda_ipc(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("HALO_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]
