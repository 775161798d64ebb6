"""GPU-vs-ORACLE parity at the sizes BASELINE.json names ("next round" item 3), plus the degenerate inputs that
used to be able to overflow the fix-up work lists.  Tolerance: |a-b| <= 1e-5*|b| + 1e-6 (helpers.py); the number of cells that
violate the pure 1e-5 relative bound of north_star is printed and bounded."""
import numpy as np
import pytest

import synth
from helpers import assert_parity, compare_layer

pytestmark = pytest.mark.gpu

KEYS = ("slope", "step", "roughness", "traversability")


def _report(tag, reports):
    for r in reports:
        print(f"{tag} {r['name']}: cells {r['cells']} bit-exact {r['bit_exact']} rel-only violations {r['rel_only_violations']} "
              f"max abs {r['max_abs']:.3g}")
        # cells beyond the pure relative bound sit where the layer value is ~0 (x within ~1 % of critical, absolute error ~1e-7) —
        # SURVEY.md C.2 measured 127 of 580 644 (1 in 4 600) for a plain fp32 path; the certified path must stay below that rate
        assert r["rel_only_violations"] <= max(4, r["cells"] // 5000), r


def test_config2_2048_both_kernels_against_oracle(te, ctx, oracle):
    """BASELINE config 2: 2048 x 2048, full chain, mixed terrain with 1 % holes; fused and literal kernels vs the oracle."""
    n = 2048
    z = synth.terrain(n, n, 0.02, 2048, "mixed")
    ref = oracle.chain(oracle.Geometry.make(n, n, 0.02), oracle.ChainParams.yaml_defaults(0), z)
    for name, kernel in (("fused", te.KERNEL_FUSED), ("literal", te.KERNEL_GENERIC)):
        ctx.set_kernel(kernel)
        got = ctx.chain_host(te.Geometry.make(n, n, 0.02), te.ChainParams.yaml_defaults(0), z)
        _report(f"2048^2 {name}", assert_parity(got, ref))
    ctx.set_kernel(te.KERNEL_AUTO)


def test_config4_batch_of_512_maps_against_oracle(te, ctx, oracle):
    """BASELINE config 4: a batch of 256 maps of 512 x 512 in one launch; 8 of the maps are checked against the oracle."""
    import torch
    n, nmaps = 512, 256
    g = te.Geometry.make(n, n, 0.02)
    p = te.ChainParams.yaml_defaults(0)
    check = (0, 1, 37, 100, 128, 201, 254, 255)
    zs = {m: synth.terrain(n, n, 0.02, 9000 + m, "mixed") for m in check}
    z = torch.empty((nmaps, n, n), dtype=torch.float32)  # [map][col][row]: column-major maps back to back
    gen = torch.Generator().manual_seed(4)
    z.normal_(0.0, 0.05, generator=gen)
    for m in check:
        z[m] = torch.from_numpy(np.ascontiguousarray(zs[m].T))
    zd = z.cuda()
    outs = [torch.empty_like(zd) for _ in range(4)]
    ctx.set_kernel(te.KERNEL_FUSED)
    ctx.chain_batched(g, p, nmaps, zd, *outs, te.MEM_DEVICE)
    ctx.synchronize()
    ctx.set_kernel(te.KERNEL_AUTO)
    og, op = oracle.Geometry.make(n, n, 0.02), oracle.ChainParams.yaml_defaults(0)
    for m in check:
        ref = oracle.chain(og, op, zs[m])
        got = {k: o[m].cpu().numpy().T for k, o in zip(KEYS, outs)}
        _report(f"batch map {m}", assert_parity(got, ref))


def test_config3_8192_against_oracle_and_eight_slabs(te, ctx, oracle):
    """BASELINE config 3: the 8192 x 8192 map (1 % holes).  The whole map is checked against the oracle (67 M cells; seconds on the
    GPU box's host cores), then recomputed as the 8 column slabs of the 8-GPU tiling (8192 x 1024 + 4 halo columns): every slab
    must be bit-identical to the whole-map result, seams included."""
    import torch
    import bench
    n = 8192
    z = bench.terrain_torch(torch, n, 0, n, n, 11, 0.01, torch.device("cuda"))  # (cols, rows) on the device
    torch.cuda.synchronize()   # the context runs on its own stream: the input must be complete
    g, p = te.Geometry.make(n, n, 0.02), te.ChainParams.yaml_defaults(0)
    outs = [torch.empty((n, n), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.set_stream(None)
    ctx.set_kernel(te.KERNEL_FUSED)
    ctx.chain(g, p, z, *outs, te.MEM_DEVICE)
    ctx.synchronize()
    cnt = ctx.flag_counters()
    assert cnt[2] == 0, cnt
    zh = z.cpu().numpy().T                                    # rows x cols view, column-major storage
    ref = oracle.chain(oracle.Geometry.make(n, n, 0.02), oracle.ChainParams.yaml_defaults(0), zh)
    got = {k: o.cpu().numpy().T for k, o in zip(KEYS, outs)}
    _report("8192^2 fused", assert_parity(got, ref))
    del ref, got
    halo = 4
    for r in range(8):
        b, e = r * 1024, (r + 1) * 1024
        hl, hr = min(halo, b), min(halo, n - e)
        part = z[b - hl:e + hr].contiguous()
        so = [torch.empty((e - b, n), dtype=torch.float32, device="cuda") for _ in range(4)]
        ctx.chain(g, p, part, *so, te.MEM_DEVICE, slab=te.Slab(b, e - b, hl, hr))
        ctx.synchronize()
        for w, o in zip(outs, so):
            assert torch.equal(w[b:e].view(torch.int32), o.view(torch.int32)), r
    ctx.set_kernel(te.KERNEL_AUTO)


def test_critical_step_that_is_not_a_float(te, ctx, oracle):
    """the reference compares (double)step_height > critical.  With critical = 0.1 (float(0.1) > 0.1) a step height of
    exactly float(0.1) counts as critical; terraces of 0 / 0.1f hit that everywhere."""
    rows, cols = 128, 120
    z = np.zeros((rows, cols), np.float32)
    z[:, 40:80] = np.float32(0.1)
    z[60:100, :] += np.float32(0.1)
    z[10:14, 10:14] = np.nan
    for crit in (0.1, 0.05, 0.2):
        pt, po = te.ChainParams.yaml_defaults(0), oracle.ChainParams.yaml_defaults(0)
        pt.step_critical = po.step_critical = crit
        ref = oracle.chain(oracle.Geometry.make(rows, cols, 0.02), po, z)
        for kernel in (te.KERNEL_FUSED, te.KERNEL_GENERIC):
            ctx.set_kernel(kernel)
            got = ctx.chain_host(te.Geometry.make(rows, cols, 0.02), pt, z)
            r = compare_layer(got["step"], ref["step"], "step")
            assert r["nan_mismatch"] == 0 and r["out_of_tol"] == 0 and r["branch_mismatch"] == 0, (crit, kernel, r)
    ctx.set_kernel(te.KERNEL_AUTO)


def test_every_cell_on_the_slow_path(te, ctx, oracle):
    """A map on which the fp32 stencil certifies nothing (an exactly planar, tilted surface riddled with holes): every cell
    lands on the tier-2 list and most go on to tier 3.  The lists are sized for the whole launch, so nothing is dropped
    (VERDICT r1 / ADVICE r1: the lists used to be capped and overflowed silently)."""
    rows, cols = 512, 384
    i, j = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    z = (0.25 * i + 0.125 * j).astype(np.float32) * np.float32(0.02)
    z[(i % 7 == 3) & (j % 5 == 1)] = np.nan
    ref = oracle.chain(oracle.Geometry.make(rows, cols, 0.02), oracle.ChainParams.yaml_defaults(0), z)
    ctx.set_kernel(te.KERNEL_FUSED)
    got = ctx.chain_host(te.Geometry.make(rows, cols, 0.02), te.ChainParams.yaml_defaults(0), z)
    cnt = ctx.flag_counters()
    ctx.set_kernel(te.KERNEL_AUTO)
    valid = int(np.isfinite(z).sum())
    assert cnt[2] == 0, cnt                       # no overflow
    assert cnt[0] >= valid, (cnt, valid)           # every valid cell was flagged
    assert cnt[1] >= cnt[0] and cnt[1] % 512 == 0  # reserved entries come in warp-private chunks
    assert_parity(got, ref)


def test_auto_falls_back_when_a_pointer_is_misaligned(te, ctx):
    """under TE_KERNEL_AUTO a launch the fused stencil cannot take (elevation not 16-byte aligned for TMA, outputs not
    8-byte aligned, only some of the normal layers) runs the generic kernel; TE_KERNEL_FUSED reports TE_ERR_UNSUPPORTED."""
    import torch
    rows, cols = 256, 192
    zh = synth.terrain(rows, cols, 0.02, 5, "mixed")
    g, p = te.Geometry.make(rows, cols, 0.02), te.ChainParams.yaml_defaults(0)
    z_al = torch.from_numpy(np.ascontiguousarray(zh.T)).reshape(-1).cuda()
    z_mis = torch.empty(rows * cols + 4, dtype=torch.float32, device="cuda")[1:rows * cols + 1]   # 4-byte offset
    z_mis.copy_(z_al)
    torch.cuda.synchronize()   # the context runs on its own stream
    new = lambda: torch.empty(rows * cols + 4, dtype=torch.float32, device="cuda")  # noqa: E731
    ctx.set_kernel(te.KERNEL_AUTO)
    base = [new()[:rows * cols] for _ in range(4)]
    ctx.chain(g, p, z_al, *base, te.MEM_DEVICE)
    l0, _ = ctx.stats()
    mis = [new()[1:rows * cols + 1] for _ in range(4)]
    ctx.chain(g, p, z_mis, *mis, te.MEM_DEVICE)   # must not fault, must not fail
    ctx.synchronize()
    for a, b in zip(base, mis):
        ok = ~torch.isnan(a)
        assert torch.equal(torch.isnan(a), torch.isnan(b))
        d = (a[ok].double() - b[ok].double()).abs()
        assert bool((d <= 1e-5 * b[ok].double().abs() + 1e-6).all())       # fused stencil vs literal kernel
    ctx.set_kernel(te.KERNEL_FUSED)
    with pytest.raises(te.TEError) as e:
        ctx.chain(g, p, z_mis, *mis, te.MEM_DEVICE)
    assert e.value.code == -4
    ctx.set_kernel(te.KERNEL_AUTO)
