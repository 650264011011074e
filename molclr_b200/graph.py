"""GraphPlan: the destination-sorted CSR (+ source-sorted transpose, per-node bond-class counts and
graph segments) built ONCE per batch on the device and shared by all layers, forward and backward.

This replaces the per-layer, per-forward ``add_self_loops`` / CPU-built ``self_loop_attr`` / H2D copy /
``cat`` of the reference (ginet_molclr.py:31-37, gcn_molclr.py:64-70) and the ``batch.max().item()``
device sync inside ``global_mean_pool``.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream

_ERR_BITS = {1: "node feature out of range (atom type must be < 119, chirality < 3; ginet_molclr.py:9-10)",
             2: "edge_index endpoint out of range", 4: "edge_attr out of range (bond type < 5, direction < 3)",
             8: "batch id out of range", 16: "a per-class node in-degree above 2048"}


def _al(n, a=4):
    return (n + a - 1) // a * a


class GraphPlan:
    """Device-resident int32 index structures for one batch (see include/molclr_b200.h: molclr_plan_build)."""

    __slots__ = ("N", "E", "G", "xpacked", "node2graph", "rowptr", "col", "eattr", "rowptr_t", "col_t", "cnt", "nbr", "nbr_t",
                 "gptr", "gperm", "status", "_buf", "_checked")

    def __init__(self, data, validate=True):
        x, ei, ea, batch = data.x, data.edge_index, data.edge_attr, data.batch
        if not x.is_cuda:
            raise RuntimeError("molclr_b200: batch tensors must be on a CUDA device (call data.to(device) first; "
                               "there is no CPU path)")
        for name, t in (("x", x), ("edge_index", ei), ("edge_attr", ea), ("batch", batch)):
            if t.dtype != torch.int64:
                raise TypeError(f"molclr_b200: data.{name} must be int64 (as produced by the reference's dataset), got {t.dtype}")
        N, E = x.size(0), ei.size(1)
        G = getattr(data, "num_graphs", None)
        if G is None:
            G = int(batch.max().item()) + 1 if N else 0       # what global_mean_pool does in the reference
        self.N, self.E, self.G = int(N), int(E), int(G)
        lib = _lib.load()
        ws_bytes = lib.molclr_plan_workspace_bytes(N, E, G)
        # one allocation, carved into 16-byte aligned int32 views
        sizes = [_al(N, 8), _al(N, 8), _al(N + 1, 8), _al(E, 8), _al((E + 3) // 4, 8), _al(N + 1, 8), _al(E, 8), _al(8 * N, 8), _al(G + 1, 8), _al(N, 8), _al(8 * N, 8), _al(8 * N, 8),
                 _al((ws_bytes + 3) // 4), 4]
        buf = torch.empty(sum(sizes), dtype=torch.int32, device=x.device)
        views, off = [], 0
        for s in sizes:
            views.append(buf[off:off + s]); off += s
        (self.xpacked, self.node2graph, self.rowptr, self.col, eattr32, self.rowptr_t, self.col_t, cnt32, self.gptr,
         self.gperm, self.nbr, self.nbr_t, ws, self.status) = views
        self.eattr = eattr32.view(torch.uint8)
        self.cnt = cnt32.view(torch.float32)
        self._buf = buf
        x, ei, ea, batch = x.contiguous(), ei.contiguous(), ea.contiguous(), batch.contiguous()
        check(lib.molclr_plan_build(ptr(x, torch.int64), ptr(ei, torch.int64), ptr(ea, torch.int64), ptr(batch, torch.int64),
                                    N, E, G, ptr(self.xpacked, torch.int32), ptr(self.node2graph, torch.int32),
                                    ptr(self.rowptr, torch.int32), ptr(self.col, torch.int32), ptr(self.eattr, torch.uint8),
                                    ptr(self.rowptr_t, torch.int32), ptr(self.col_t, torch.int32), ptr(self.cnt), ptr(self.nbr, torch.int32), ptr(self.nbr_t, torch.int32),
                                    ptr(self.gptr, torch.int32), ptr(self.gperm, torch.int32), ptr(ws, torch.int32), ws_bytes,
                                    ptr(self.status, torch.int32), stream()), "plan_build")
        self._checked = False
        if validate:
            self.check()

    def check(self):
        """Raises IndexError for out-of-range inputs (one 16-byte D2H read that drains the stream; the reference's embedding
        lookups raise the same way on CPU)."""
        if not self._checked:
            bits = int(self.status[0].item())
            self._checked = True
            _raise_for(bits)
        return self

    def check_deferred(self):
        """The same check without stalling the stream: the status word is copied to pinned host memory behind the plan kernels
        and examined by a later ``poll_checks`` (the next ``get_plan`` call, or ``poll_checks(block=True)``).  Safe because the
        plan kernels sanitise what they flag (out-of-range features read row 0, bad edges are dropped): the step that consumed
        a bad batch computes on the sanitised graph, and the IndexError is raised one call late instead of never."""
        if not self._checked:
            _enqueue(self)
            self._checked = True
        return self


def _raise_for(bits):
    if bits:
        raise IndexError("molclr_b200: invalid batch: " + "; ".join(m for b, m in _ERR_BITS.items() if bits & b))


_RING = 16
_slots = {}          # device index -> {"host": pinned int32 [_RING][4], "events": [...], "next": int, "pending": [slot, ...]}


def _enqueue(plan):
    dev = plan.status.device
    st = _slots.get(dev.index)
    if st is None:
        st = _slots[dev.index] = {"host": torch.zeros(_RING, 4, dtype=torch.int32).pin_memory(),
                                  "events": [torch.cuda.Event() for _ in range(_RING)], "next": 0, "pending": []}
    if len(st["pending"]) == _RING:
        poll_checks(block=True, device=dev)
    slot = st["next"]
    st["next"] = (slot + 1) % _RING
    st["host"][slot].copy_(plan.status, non_blocking=True)
    st["events"][slot].record(torch.cuda.current_stream(dev))
    st["pending"].append(slot)


def poll_checks(block=False, device=None):
    """Examines the deferred batch checks whose status words have arrived (all of them if ``block``); raises IndexError for the
    first invalid batch found."""
    for idx, st in list(_slots.items()):
        if device is not None and device.index != idx:
            continue
        while st["pending"]:
            slot = st["pending"][0]
            ev = st["events"][slot]
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            st["pending"].pop(0)
            _raise_for(int(st["host"][slot, 0]))


def get_plan(data, validate="deferred"):
    """Returns the cached plan of a batch object, building it on first use.  validate: "deferred" (default: no device sync, see
    ``GraphPlan.check_deferred``), True / "sync" (raise here, one stream drain), False (no check)."""
    plan = getattr(data, "_molclr_plan", None)
    if plan is None or plan.xpacked.device != data.x.device:
        if validate == "deferred":
            poll_checks(device=data.x.device)
        plan = GraphPlan(data, validate=validate in (True, "sync"))
        if validate == "deferred":
            plan.check_deferred()
        try:
            data._molclr_plan = plan
        except AttributeError:
            pass
    return plan
