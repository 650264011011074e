"""Whole-pass calls: the GIN encoder (and the projection head) forward / backward as ONE C call each (csrc/gin_step.cu) instead
of one Python -> C call per kernel.  Same kernels, same order, same results; what changes is the host time per step (at 512
pairs per step the Python call overhead, not the GPU, used to set the step time)."""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import GinLayer, GinModel, PlanView, check, stream


def plan_view(plan):
    pv = PlanView()
    pv.N, pv.E, pv.G = plan.N, plan.E, plan.G
    for name in ("xpacked", "node2graph", "rowptr", "col", "eattr", "rowptr_t", "col_t", "cnt", "nbr", "nbr_t", "gptr", "gperm"):
        setattr(pv, name, getattr(plan, name).data_ptr())
    return pv


def _dp(t):
    return None if t is None else t.data_ptr()


def model_struct(m, comp, with_head=True):
    """(GinModel, keepalive): device pointers of the parameters and of the weight shadows of the LAST ``m._refresh_weights``.
    Cached while the shadow buffers (``_RoundedWeights.generation``) and the BatchNorm buffers stay where they are."""
    rw = m._rounded
    key = (getattr(rw, "generation", None), int(comp), bool(with_head), tuple(p.data_ptr() for p in m._params()),
           tuple((bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), bn.momentum, bn.eps) for bn in m.batch_norms))
    hit = m.__dict__.get("_native_struct")
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    gm, keep = _build_model_struct(m, comp, with_head)
    m.__dict__["_native_struct"] = (key, gm, keep)
    return gm, keep


def _build_model_struct(m, comp, with_head):
    rw = m._rounded
    L = m.num_layer
    layers = (GinLayer * L)()
    keep = [layers, rw._cur]
    gm = GinModel()
    gm.num_layer, gm.emb_dim, gm.feat_dim = L, m.emb_dim, m.feat_dim
    gm.x_emb1, gm.x_emb2 = m.x_embedding1.weight.data_ptr(), m.x_embedding2.weight.data_ptr()
    for l in range(L):
        g, bn, ly = m.gnns[l], m.batch_norms[l], layers[l]
        w1, w2 = g.mlp[0].weight, g.mlp[2].weight
        ly.w1_hi, ly.w2_hi = rw.get(w1)[0].data_ptr(), rw.get(w2)[0].data_ptr()
        ly.w1_hi_t, ly.w2_hi_t = _dp(rw.hi_t(w1)), _dp(rw.hi_t(w2))
        if comp:
            s1, s2 = rw.b16(w1), rw.b16(w2)          # bf16 correction tiles (comp 1) or the fp16 halves (comp 2: no raw copy)
            ly.w1_b16, ly.w2_b16 = s1.data_ptr(), s2.data_ptr()
            if int(comp) != 2:
                ly.w1_raw, ly.w2_raw = rw.raw(w1).data_ptr(), rw.raw(w2).data_ptr()
            gm.w1_ld16, gm.w1_rows16, gm.w2_ld16, gm.w2_rows16 = s1.stride(1), s1.shape[1], s2.stride(1), s2.shape[1]
        ly.b1, ly.b2 = g.mlp[0].bias.data_ptr(), g.mlp[2].bias.data_ptr()
        ly.bond_type, ly.bond_dir = g.edge_embedding1.weight.data_ptr(), g.edge_embedding2.weight.data_ptr()
        ly.gamma, ly.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
        ly.running_mean, ly.running_var, ly.num_batches_tracked = bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr()
        ly.momentum, ly.eps = (0.1 if bn.momentum is None else bn.momentum), bn.eps
    gm.layers = layers
    st = m.__dict__.get("_fp16_status")
    if int(comp) == 2 and st is not None:
        gm.status = st["dev"].data_ptr()
        keep.append(st["dev"])
    if with_head:
        (wf, wfl), (w0, w0l), (w2, w2l) = rw.get(m.feat_lin.weight), rw.get(m.out_lin[0].weight), rw.get(m.out_lin[2].weight)
        gm.wf_hi, gm.wf_lo, gm.bf = wf.data_ptr(), _dp(wfl), m.feat_lin.bias.data_ptr()
        gm.w0_hi, gm.w0_lo, gm.b0 = w0.data_ptr(), _dp(w0l), m.out_lin[0].bias.data_ptr()
        gm.w2_hi, gm.w2_lo, gm.b2 = w2.data_ptr(), _dp(w2l), m.out_lin[2].bias.data_ptr()
    return gm, keep


class EncoderContext:
    """What one encoder forward leaves for its backward: the ctx buffer, the model / plan structs and the dropout seeds."""
    __slots__ = ("gm", "keep", "pv", "plan", "ctx", "comp", "training", "pool_mode", "seeds", "drop_p", "p", "p_lo", "dev")


def encoder_forward(m, plan, comp, training, pool_mode, with_head=True):
    """Node embedding -> L x (aggregate, MLP, BatchNorm statistics) -> pooled graph vectors (ginet_molclr.py:103-113).
    Returns an EncoderContext whose ``p`` / ``p_lo`` are the pooled operand pair (views into the ctx buffer)."""
    lib = _lib.load()
    e = EncoderContext()
    e.gm, e.keep = model_struct(m, comp, with_head)
    e.pv, e.plan, e.comp, e.training, e.pool_mode = plan_view(plan), plan, int(comp), int(training), int(pool_mode)
    e.dev = m.x_embedding1.weight.device
    drops = m._dropout_seeds()
    e.drop_p = float(drops[0][1])
    e.seeds = (C.c_uint32 * m.num_layer)(*[int(s) for s, _ in drops]) if e.drop_p > 0 else None
    gmp = C.byref(e.gm)
    nbytes = lib.molclr_gin_ctx_bytes(gmp, plan.N, plan.G, e.comp, e.pool_mode)
    e.ctx = torch.empty(nbytes // 4, dtype=torch.float32, device=e.dev)
    sbytes = lib.molclr_gin_scratch_bytes(gmp, plan.N, plan.G, 0)
    scratch = _scratch(e.dev, sbytes)
    check(lib.molclr_gin_encoder_fwd(gmp, C.byref(e.pv), e.comp, e.training, e.pool_mode, e.seeds, e.drop_p, e.ctx.data_ptr(), nbytes,
                                     scratch.data_ptr(), sbytes, stream()), "gin_encoder_fwd")
    pp, pl, ld = C.c_void_p(), C.c_void_p(), C.c_int64()
    lib.molclr_gin_ctx_pooled(gmp, plan.N, plan.G, e.comp, e.pool_mode, e.ctx.data_ptr(), C.byref(pp), C.byref(pl), C.byref(ld))
    view = lambda a: e.ctx[(a - e.ctx.data_ptr()) // 4:(a - e.ctx.data_ptr()) // 4 + plan.G * ld.value].view(plan.G, ld.value)[:, :m.emb_dim]
    e.p = view(pp.value)
    e.p_lo = view(pl.value) if pl.value else None
    return e


_SCRATCH = {}


def _scratch(dev, nbytes):
    """One scratch buffer per device, grown on demand (temporaries of a pass; free again when the call's kernels have run --
    all passes of a process are ordered on one stream)."""
    key = (dev.type, dev.index, stream())
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() * 4 < nbytes:
        buf = _SCRATCH[key] = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
    return buf


def grad_views(m, e, flat, first, count):
    lib = _lib.load()
    n = 2 + 8 * m.num_layer + 6
    offs = (C.c_int64 * n)()
    total = lib.molclr_gin_grad_layout(C.byref(e.gm), offs)
    shapes = _grad_shapes(m)
    return [flat[offs[k]:offs[k] + _numel(shapes[k])].view(shapes[k]) for k in range(first, first + count)], total


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def _grad_shapes(m):
    D, H, F = m.emb_dim, 2 * m.emb_dim, m.feat_dim
    shapes = [(119, D), (3, D)]
    for _ in range(m.num_layer):
        shapes += [(H, D), (H,), (D, H), (D,), (5, D), (3, D), (D,), (D,)]
    shapes += [(F, D), (F,), (F, F), (F,), (F // 2, F), (F // 2,)]
    return shapes


def grad_buffer(m, e):
    total = _lib.load().molclr_gin_grad_layout(C.byref(e.gm), None)
    return torch.empty(total, dtype=torch.float32, device=e.dev)


def encoder_backward(m, e, g_p, flat, ordered, on_layer_done=None):
    """Backward of ``encoder_forward``: fills the 2 + 8 L encoder slots of ``flat``.  g_p None: taken from the scratch slot the head
    backward of the same context just wrote.  on_layer_done(l): called from inside the C call when the gradients of layer l
    (l = L-1 .. 0; -1 = the node-embedding tables) have been enqueued."""
    lib = _lib.load()
    sbytes = lib.molclr_gin_scratch_bytes(C.byref(e.gm), e.plan.N, e.plan.G, int(ordered))
    scratch = _scratch(e.dev, sbytes)
    gp = None if g_p is None else ops.ptr2d(g_p.contiguous())
    errs = []

    def _cb(layer, _user):
        try:
            on_layer_done(layer)
        except BaseException as ex:          # an exception must not unwind through the C frame
            errs.append(ex)
    cb = _lib.LAYER_CB(_cb) if on_layer_done is not None else _lib.LAYER_CB()
    check(lib.molclr_gin_encoder_bwd(C.byref(e.gm), C.byref(e.pv), e.comp, e.training, e.pool_mode, e.seeds, e.drop_p, e.ctx.data_ptr(), gp,
                                     int(ordered), flat.data_ptr(), scratch.data_ptr(), sbytes, cb, None, stream()), "gin_encoder_bwd")
    if errs:
        raise errs[0]


def grad_slices(m, e):
    """Float ranges of the flat gradient buffer: {'embed': (lo, hi), 'layer': [(lo, hi)] * L, 'head': (lo, hi)}."""
    n = 2 + 8 * m.num_layer + 6
    offs = (C.c_int64 * n)()
    total = _lib.load().molclr_gin_grad_layout(C.byref(e.gm), offs)
    L = m.num_layer
    bounds = list(offs) + [total]
    return {"embed": (bounds[0], bounds[2]), "layer": [(bounds[2 + 8 * l], bounds[2 + 8 * (l + 1)]) for l in range(L)],
            "head": (bounds[2 + 8 * L], total), "total": total}


def proj_head_forward(m, e):
    lib = _lib.load()
    G, F = e.plan.G, m.feat_dim
    h = torch.empty(G, F, device=e.dev)
    out = torch.empty(G, F // 2, device=e.dev)
    check(lib.molclr_proj_head_fwd(C.byref(e.gm), e.plan.N, G, e.comp, e.pool_mode, e.ctx.data_ptr(), h.data_ptr(), out.data_ptr(), stream()),
          "proj_head_fwd")
    return h, out


def proj_head_backward(m, e, g_h, g_out, flat, ordered):
    lib = _lib.load()
    sbytes = lib.molclr_gin_scratch_bytes(C.byref(e.gm), e.plan.N, e.plan.G, int(ordered))
    scratch = _scratch(e.dev, sbytes)
    gh = None if g_h is None else ops.ptr(g_h.contiguous())
    check(lib.molclr_proj_head_bwd(C.byref(e.gm), e.plan.N, e.plan.G, e.comp, e.pool_mode, e.ctx.data_ptr(), gh, ops.ptr(g_out.contiguous()),
                                   int(ordered), flat.data_ptr(), scratch.data_ptr(), sbytes, stream()), "proj_head_bwd")


def add_inplace(y, x):
    check(_lib.load().molclr_add_inplace(y.data_ptr(), x.data_ptr(), y.numel(), stream()), "add_inplace")
    return y
