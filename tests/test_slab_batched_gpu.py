"""Column slabs and batches through the C ABI with device pointers (torch only owns the memory)."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()  # (cols, rows) tensor == column-major layer


def test_slabs_are_bit_identical_to_the_whole_map(te, ctx):
    import torch
    rows, cols = 192, 250
    z = synth.terrain(rows, cols, 0.02, 31, "mixed")
    g = te.Geometry.make(rows, cols, 0.02)
    p = te.ChainParams.yaml_defaults(0)
    ctx.set_kernel(te.KERNEL_AUTO)
    ctx.set_stream(None)
    zd = _dev(torch, z)
    whole = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g, p, zd, *whole, te.MEM_DEVICE)
    ctx.synchronize()
    for cuts in ((0, 100, 250), (0, 7, 130, 131, 250)):
        for b, e in zip(cuts, cuts[1:]):
            hl, hr = min(4, b), min(4, cols - e)
            slab = te.Slab(b, e - b, hl, hr)
            part_in = zd[b - hl:e + hr].contiguous()
            outs = [torch.empty((e - b, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
            ctx.chain(g, p, part_in, *outs, te.MEM_DEVICE, slab=slab)
            ctx.synchronize()
            for w, o in zip(whole, outs):
                assert torch.equal(w[b:e].view(torch.int32), o.view(torch.int32)), (b, e)
    with pytest.raises(te.TEError):  # a halo smaller than the dependency radius is refused, not silently wrong
        ctx.chain(g, p, zd[98:250].contiguous(), *[torch.empty((150, rows), dtype=torch.float32, device="cuda") for _ in range(4)],
                  te.MEM_DEVICE, slab=te.Slab(100, 150, 2, 0))


def test_batched_equals_one_by_one(te, ctx):
    import torch
    rows, cols, n = 128, 96, 5
    g = te.Geometry.make(rows, cols, 0.02)
    p = te.ChainParams.yaml_defaults(0)
    maps = np.stack([np.ascontiguousarray(synth.terrain(rows, cols, 0.02, 1000 + k, "mixed").T) for k in range(n)])
    zd = torch.from_numpy(maps).cuda()
    outs = [torch.empty((n, cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.set_stream(None)
    ctx.chain_batched(g, p, n, zd, *outs, te.MEM_DEVICE)
    ctx.synchronize()
    for k in range(n):
        one = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
        ctx.chain(g, p, zd[k], *one, te.MEM_DEVICE)
        ctx.synchronize()
        for a, b in zip(outs, one):
            assert torch.equal(a[k].view(torch.int32), b.view(torch.int32))


def test_size_independent_properties_at_scale(te, ctx):
    """At a size the oracle cannot check in seconds: idempotence, NaN pattern and range invariants."""
    import torch
    import bench
    rows = cols = 4096
    z = bench.terrain_torch(torch, rows, 0, cols, cols, 7, 0.01, torch.device("cuda"))
    torch.cuda.synchronize()   # the context runs on its own stream: the input must be complete
    g = te.Geometry.make(rows, cols, 0.02)
    p = te.ChainParams.yaml_defaults(0)
    ctx.set_stream(None)
    a = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    b = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g, p, z, *a, te.MEM_DEVICE)
    ctx.chain(g, p, z, *b, te.MEM_DEVICE)
    ctx.synchronize()
    hole = ~torch.isfinite(z)
    for x, y in zip(a, b):
        assert torch.equal(x.view(torch.int32), y.view(torch.int32))       # deterministic
        v = x[torch.isfinite(x)]
        assert float(v.min()) >= 0.0 and float(v.max()) <= 1.0
    slope, step, rough, trav = a
    assert torch.equal(torch.isnan(slope), hole) and torch.equal(torch.isnan(rough), hole) and torch.equal(torch.isnan(trav), hole)
    assert int(torch.isnan(step).sum()) < int(hole.sum())                  # step is defined inside small holes too
    w = torch.tensor(np.float32(1.0) / np.float32(3.0), device="cuda")
    ok = ~hole
    assert torch.equal(trav[ok], (w * ((slope[ok] + step[ok]) + rough[ok])))  # the fuse is exactly float32 left-to-right
    # translation of the map position must not change anything but the on-circle memberships: same NaN pattern
    g2 = te.Geometry.make(rows, cols, 0.02, (123.456, -78.9))
    c = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g2, p, z, *c, te.MEM_DEVICE)
    ctx.synchronize()
    assert torch.equal(torch.isnan(c[0]), hole)
    assert torch.equal(c[0][ok & torch.isfinite(c[0])], slope[ok & torch.isfinite(c[0])])  # slope/roughness windows have no on-circle offsets


def test_pipelined_host_path_equals_device_path(te, ctx):
    """te_chain(TE_MEM_HOST) on a large map runs as overlapped column chunks; it must equal the one-shot device result bitwise."""
    import torch
    import bench
    rows, cols = 2048, 2304
    z = bench.terrain_torch(torch, rows, 0, cols, cols, 9, 0.01, torch.device("cuda"))
    torch.cuda.synchronize()   # the context runs on its own stream: the input must be complete
    g = te.Geometry.make(rows, cols, 0.02)
    p = te.ChainParams.yaml_defaults(0)
    ctx.set_stream(None)
    dev = [torch.empty((cols, rows), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.chain(g, p, z, *dev, te.MEM_DEVICE)
    ctx.synchronize()
    h_in = z.cpu().pin_memory()
    h_out = [torch.empty((cols, rows), dtype=torch.float32).pin_memory() for _ in range(4)]
    ctx.chain(g, p, h_in.data_ptr(), *[o.data_ptr() for o in h_out], te.MEM_HOST)
    for a, b in zip(dev, h_out):
        assert torch.equal(a.cpu().view(torch.int32), b.view(torch.int32))
