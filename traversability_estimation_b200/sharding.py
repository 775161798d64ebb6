"""Column-slab tiling of one map over the ranks of a torch.distributed job (one process per GPU).

The chain is a pure stencil of `elevation`, so a map larger than one GPU is split along the COLUMN
index (the slow storage axis of the column-major layers: a slab is one contiguous block) and the
only exchange step is a one-shot swap of `halo` boundary columns of `elevation` with the two
neighbours (NCCL send/recv over NVLink; gloo in the CPU tests).  Intermediates are recomputed in
the halo, outputs stay sharded.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class SlabPlan:
    rank: int
    world: int
    cols_total: int
    col_begin: int
    col_count: int
    halo_left: int
    halo_right: int

    @property
    def buffer_cols(self) -> int:
        return self.halo_left + self.col_count + self.halo_right


def plan_slab(cols_total: int, world: int, rank: int, halo: int) -> SlabPlan:
    """Contiguous, near-equal column ranges; halos clipped at the map edges.

    Every rank must own at least `halo` columns: the exchange is one hop (a halo comes from the direct neighbour's owned
    columns only), so narrower slabs would need data from two ranks away."""
    if not (0 <= rank < world) or cols_total < world:
        raise ValueError((cols_total, world, rank))
    base, extra = divmod(cols_total, world)
    if world > 1 and base < halo:
        raise ValueError(f"{cols_total} columns over {world} ranks leaves slabs of {base} columns, narrower than the "
                         f"halo of {halo}: use fewer ranks (the halo exchange is one hop)")
    begin = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    hl = min(halo, begin)
    hr = min(halo, cols_total - (begin + count))
    return SlabPlan(rank, world, cols_total, begin, count, hl, hr)


def exchange_halo(dist, buf, plan: SlabPlan, halo: int):
    """buf: tensor of shape (buffer_cols, rows) holding [left halo | owned columns | right halo]; fills both
    halos from the neighbours' owned boundary columns.  All ranks must call it."""
    if plan.world == 1:
        return
    hl, hr, n = plan.halo_left, plan.halo_right, plan.col_count
    ops = []
    if plan.rank > 0:
        ops.append(dist.P2POp(dist.isend, buf[hl:hl + min(halo, n)], plan.rank - 1))
        ops.append(dist.P2POp(dist.irecv, buf[0:hl], plan.rank - 1))
    if plan.rank < plan.world - 1:
        ops.append(dist.P2POp(dist.isend, buf[hl + n - min(halo, n):hl + n], plan.rank + 1))
        ops.append(dist.P2POp(dist.irecv, buf[hl + n:hl + n + hr], plan.rank + 1))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class PeerHalo:
    """Peer-mapped halo exchange of one slab buffer through the C ABI (te_halo_pull): every rank exports its buffer and a
    "layer ready" event once (CUDA IPC), opens its neighbours', and from then on a halo exchange is two asynchronous
    device-to-device copies straight out of the neighbours' buffers over NVLink, queued on the context stream.
    torch.distributed only carries the 64-byte handles at set-up."""

    def __init__(self, dist, ctx, te, buf, plan: SlabPlan):
        self.ctx, self.te, self.plan, self.buf = ctx, te, plan, buf
        self.slab = te.Slab(plan.col_begin, plan.col_count, plan.halo_left, plan.halo_right)
        self.ready, ev_handle = ctx.event_create_ipc()
        mine = {"mem": ctx.ipc_export(buf.data_ptr()), "ev": ev_handle,
                "slab": (plan.col_begin, plan.col_count, plan.halo_left, plan.halo_right)}
        everyone = [None] * plan.world
        dist.all_gather_object(everyone, mine)
        self._opened = []
        self.left = self._open(everyone[plan.rank - 1]) if plan.rank > 0 else None
        self.right = self._open(everyone[plan.rank + 1]) if plan.rank < plan.world - 1 else None

    def _open(self, info):
        ptr = self.ctx.ipc_open(info["mem"])
        ev = self.ctx.event_open_ipc(info["ev"])
        self._opened.append((ptr, ev))
        return self.te.HaloPeer(ptr, self.te.Slab(*info["slab"]), ev)

    def publish(self):
        """This rank's owned columns are valid (as far as the context stream is concerned)."""
        self.ctx.event_record(self.ready)

    def pull(self, g):
        self.ctx.halo_pull(g, self.slab, self.buf, self.left, self.right)

    def close(self):
        for ptr, ev in self._opened:
            self.ctx.event_destroy(ev)
            self.ctx.ipc_close(ptr)
        self._opened = []
        self.ctx.event_destroy(self.ready)
