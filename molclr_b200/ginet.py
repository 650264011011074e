"""GINet: drop-in for ``models/ginet_molclr.py`` (same constructor, ``forward(data) -> (h, out)``,
``state_dict`` keys) whose forward AND backward run entirely on the hand-written sm_100a kernels.

Per view the kernel sequence is (N nodes, D=emb_dim, H=2D):

    plan (once per batch)      CSR + transpose + bond-class counts + graph segments
    embed_nodes_fwd            h0 = E1[x0] + E2[x1]                                   ginet_molclr.py:103
    per layer l:
      gine_aggregate_fwd       a_l = sum_j f(z_{l-1})[j] + bond table, self loop last   :29-44  (f = BN+ReLU of l-1, fused)
      gemm (bias, ReLU)        u_l = relu(a_l W1^T + b1)                               :19-23,46-47
      gemm (bias, tile stats)  z_l = u_l W2^T + b2, column mean/M2 per 128-row tile
      bn_fwd_finalize          batch statistics -> (scale, shift, mean, invstd); running stats   :107
    pool_fwd                   p = mean_g( BN_L(z_L) )                                 :113
    gemm x3                    h = feat_lin(p); out = out_lin(h)                       :114-115

and the backward mirrors it.  Both sequences are issued from C in ONE call each (``molclr_gin_encoder_fwd/bwd``,
``molclr_proj_head_fwd/bwd`` in csrc/gin_step.cu, driven by ``molclr_b200/native.py``); the per-kernel wrappers of ``ops.py`` remain
for the GCN path, the fine-tune heads and the tests.  Nothing here falls back to PyTorch operators.
"""
import os

import torch
from torch import nn

from . import _lib, native, ops
from .graph import get_plan

num_atom_type = 119      # including the extra mask token   (ginet_molclr.py:9)
num_chirality_tag = 3
num_bond_type = 5        # including aromatic and self-loop  (ginet_molclr.py:12)
num_bond_direction = 3


class GINEConv(nn.Module):
    """Parameter container with the reference's names (ginet_molclr.py:16-27); the computation lives in
    the fused kernels driven by ``GINet``."""

    def __init__(self, emb_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(emb_dim, 2 * emb_dim), nn.ReLU(), nn.Linear(2 * emb_dim, emb_dim))
        self.edge_embedding1 = nn.Embedding(num_bond_type, emb_dim)
        self.edge_embedding2 = nn.Embedding(num_bond_direction, emb_dim)
        nn.init.xavier_uniform_(self.edge_embedding1.weight.data)
        nn.init.xavier_uniform_(self.edge_embedding2.weight.data)


class _RoundedWeights:
    """Tensor-core operand forms of the GEMM weights -- hi = tf32(w), lo = tf32(w - hi), an unrounded copy with 128-byte rows
    (of the transpose, for weights stored [in, out]) and its bf16 correction tiles -- derived from the CURRENT parameter values
    by ONE kernel launch at the start of every forward (``refresh``).  No VALUE is cached across forwards: in-place parameter
    updates (``torch.optim.Adam(fused=True)``, ``param.data`` writes, ``dist.broadcast``) do not bump tensor version counters,
    so any cache keyed on them serves stale weights.  (The shadow buffers themselves are reused.)"""

    def __init__(self):
        self._cur = {}
        self._key, self._launch = None, None

    def refresh(self, specs):
        """specs: [(parameter, ops.W_* flags)].  The shadows of exactly these parameters are valid until the next refresh.  The
        shadow BUFFERS are kept while the parameter set (objects, storage, flags) stays the same, so a refresh is one kernel
        launch with a prepared descriptor table; the VALUES are re-derived every time."""
        key = tuple((id(p), p.data_ptr(), f) for p, f in specs)
        if key != self._key:
            outs, self._launch = ops.prepare_weights([(p.detach(), f) for p, f in specs], want_relaunch=True)
            self._cur = {id(p): (p.data_ptr(), o) for (p, _), o in zip(specs, outs)}
            self._key = key
            self.generation = getattr(self, "generation", 0) + 1
        else:
            self._launch()

    def _entry(self, p, key):
        hit = self._cur.get(id(p))
        if hit is not None and hit[0] == p.data_ptr() and hit[1][key] is not None:
            return hit[1]
        return None

    def get(self, p):
        e = self._entry(p, "hi")
        if e is not None:
            return e["hi"], e["lo"]
        return ops.split_tf32(p.detach())                 # not part of the last refresh: derived now, never cached

    def raw(self, p, transpose=False):
        """Unrounded copy (of the transpose, for weights stored [in, out]) with 128-byte aligned rows: the K-major B operand of
        the compensated GEMM."""
        e = self._entry(p, "raw")
        if e is not None:
            return e["raw"]
        return ops.prepare_weights([(p.detach(), ops.W_RAW_T if transpose else ops.W_RAW)])[0]["raw"]

    def hi_t(self, p):
        """tf32(p^T) with 128-byte rows (None unless part of the last refresh)."""
        e = self._entry(p, "hi_t")
        return None if e is None else e["hi_t"]

    def b16(self, p):
        """bf16 correction tiles [2, rows16, ld16] of ``raw(p)`` (None unless part of the last refresh)."""
        e = self._entry(p, "b16")
        return None if e is None else e["b16"]


PRECISIONS = ("fp16x3", "tf32x3", "tf32")


class _EncoderBase(nn.Module):
    """Shared plumbing of the GINet / GCN drop-ins.

    ``precision`` (attribute, not a constructor argument -- the constructor is the reference's):
      * ``"fp16x3"`` (default) and ``"tf32x3"``: every FORWARD contraction is an error-compensated three-product form
        (~fp32 accuracy), so pre-activations -- and with them the ReLU masks the backward pass depends on --
        match the fp32 reference; BACKWARD contractions are single-pass TF32.  ``"tf32x3"``: a TF32 pass plus two bf16
        correction passes.  ``"fp16x3"``: the GINEConv MLP / GCNConv products in the fp16 three-product form instead
        (``molclr_gemm_args.compensate = 2``: operands split into two fp16 halves, 22 significand bits): a smaller
        rounding error than ``"tf32x3"`` at 3/4 of its tensor work and 2/3 of its operand traffic, valid while the
        activations stay inside fp16's range -- |a| <= 65504; a violation is detected on the device and raised as
        ``FloatingPointError`` by the next forward (BatchNorm keeps the activations of this network at O(1));
        the projection / prediction heads keep the TF32 form.
      * ``"tf32"``: single-pass TF32 everywhere (fastest; activations carry ~1e-3 relative error and the
        resulting ReLU mask flips show up as percent-level noise in gradients).
    """
    precision = os.environ.get("MOLCLR_B200_PRECISION", "fp16x3")      # class default; the environment variable overrides it process-wide

    def _check_input(self, data):
        if not 0 <= self.drop_ratio < 1:
            raise ValueError(f"dropout probability has to be between 0 and 1, but got {self.drop_ratio}")
        if self.pool_name not in ops.POOL_MODES:
            # the reference leaves self.pool unset for unknown names and fails at forward (ginet_molclr.py:83-88,113)
            raise AttributeError(f"'{type(self).__name__}' object has no attribute 'pool'")

    deterministic = False      # True: weight gradients are summed in a fixed order (bit-reproducible run to run); see DESIGN.md

    # per-instance caches of device state (weight-shadow buffers, the native model struct, the fp16 status word with its stream and
    # event): rebuilt on demand, never copied or pickled with the module (copy.deepcopy(model), torch.save(model))
    _TRANSIENT = ("_gemm_weight_specs", "_native_struct", "_fp16_status")

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items() if k not in self._TRANSIENT}
        if "_rounded" in state:
            state["_rounded"] = _RoundedWeights()
        return state

    def _gemm_weights(self, comp):
        """(parameter, ops.W_* operand forms) of every contraction of the model.  GINEConv MLP weights / GCNConv weights: hi
        (backward dX products, single-pass forward) and, for the compensated forward, the unrounded K-major copy + its bf16
        correction tiles; every other nn.Linear (projection / prediction heads): hi (+ lo for the explicit 3-pass product)."""
        h3 = int(comp) == 2      # fp16 three-product form: the fp16 halves of the weight instead of the raw copy + bf16 correction tiles
        enc = ops.W_HI | ((ops.W_H16 if h3 else ops.W_B16) if comp else 0)
        head = ops.W_HI | (ops.W_LO if comp else 0)
        raw = ops.W_RAW if (comp and not h3) else 0
        specs, seen = [], set()
        for g in self.gnns:
            if hasattr(g, "mlp"):                         # (+ the transposed tf32 copies: K-major operands of the backward dX products)
                ws = [(g.mlp[0].weight, enc | ops.W_HI_T | raw), (g.mlp[2].weight, enc | ops.W_HI_T | raw)]
            else:                                         # GCNConv: stored [in, out]
                ws = [(g.weight, ops.W_HI | ((ops.W_H16 | ops.W_T16) if h3 else (ops.W_B16 | ops.W_RAW_T) if comp else 0))]
            specs += ws
            seen.update(id(w) for w, _ in ws)
        for mod in self.modules():
            if isinstance(mod, nn.Linear) and id(mod.weight) not in seen:
                seen.add(id(mod.weight))
                specs.append((mod.weight, head))
        return specs

    def _refresh_weights(self, comp):
        cache = self.__dict__.setdefault("_gemm_weight_specs", {})
        specs = cache.get(int(comp))
        if specs is None:
            specs = cache[int(comp)] = self._gemm_weights(comp)
        self._rounded.refresh(specs)
        if int(comp) == 2:
            self._fp16_range_check()

    fp16_check_every = 8       # forwards between two looks at the status word (1: every forward)

    def check_fp16_range(self):
        """Blocking form of the range check of ``precision = "fp16x3"`` (drains the stream): raises ``FloatingPointError`` if any forward
        product since the last check saw an activation beyond fp16's finite range.  The trainer calls it once per epoch."""
        st = self.__dict__.get("_fp16_status")
        if st is None:
            return
        st["ev"] = None
        if int(st["dev"].item()) & _lib.STATUS_FP16_RANGE:
            st["dev"].zero_()
            raise FloatingPointError("molclr_b200: an activation exceeded fp16's finite range (65504) in a forward product of precision "
                                     "'fp16x3' (the value was clamped: results since the last check are wrong); use model.precision = 'tf32x3'")

    def _fp16_range_check(self):
        """``precision = "fp16x3"``: the forward products report an activation beyond fp16's finite range through a sticky device
        word (``molclr_gin_model.status``).  Every ``fp16_check_every``-th forward copies it to pinned memory on a side stream and a
        later forward examines the copy -- no stream drain; the error arrives a few calls late."""
        st = self.__dict__.get("_fp16_status")
        dev = next(self.parameters()).device
        if st is None or st["dev"].device != dev:            # (first use, or the model has moved to another device)
            st = self.__dict__["_fp16_status"] = {"dev": torch.zeros(1, dtype=torch.int32, device=dev),
                                                  "host": torch.zeros(1, dtype=torch.int32).pin_memory(), "ev": None, "calls": 0}
        st["calls"] += 1
        if st["ev"] is not None and st["ev"].query():          # the last copy has arrived: examine it
            st["ev"] = None
            if int(st["host"][0]) & _lib.STATUS_FP16_RANGE:
                st["dev"].zero_()
                raise FloatingPointError("molclr_b200: an activation exceeded fp16's finite range (65504) in a forward product of "
                                         "precision 'fp16x3' (the value was clamped: results since the last check are wrong); "
                                         "use model.precision = 'tf32x3'")
        # a new copy on forwards 1, 1 + every, 1 + 2 every, ... (the device word is sticky: a later look still finds a violation; the
        # first forward takes one, so that the side stream's one-time set-up cost falls into the warm-up of a run)
        if st["ev"] is None and (st["calls"] - 1) % self.fp16_check_every == 0:
            # the 4-byte copy runs on a side stream behind everything enqueued so far: on the compute stream it would queue behind any
            # host-to-device batch copy in flight on the same copy engine (measured: 0.25 ms per step in bench.py's e2e loop)
            side = st.get("stream")
            if side is None:
                side = st["stream"] = torch.cuda.Stream(device=st["dev"].device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                st["host"].copy_(st["dev"], non_blocking=True)
                st["ev"] = torch.cuda.Event()
                st["ev"].record(side)

    def _dropout_seeds(self):
        """One counter-hash seed per layer and forward call (drawn from torch's CPU generator, so torch.manual_seed makes
        a run reproducible); (0, 0.0) entries when dropout is inactive (eval mode or drop_ratio 0)."""
        if not self.training or self.drop_ratio <= 0:
            return [(0, 0.0)] * self.num_layer
        base = torch.randint(0, 2 ** 31 - 1, (self.num_layer,))
        return [(int(b), float(self.drop_ratio)) for b in base]


class GINet(_EncoderBase):
    """ginet_molclr.py:50-117.

    Args:
        num_layer (int): the number of GNN layers
        emb_dim (int): dimensionality of embeddings
        feat_dim (int): dimensionality of the returned representation ``h``
        drop_ratio (float): dropout rate
        pool (str): 'mean' | 'add' | 'max'
    """

    def __init__(self, num_layer=5, emb_dim=300, feat_dim=256, drop_ratio=0, pool="mean"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio = num_layer, emb_dim, feat_dim, drop_ratio
        self.pool_name = pool
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        self.out_lin = nn.Sequential(nn.Linear(feat_dim, feat_dim), nn.ReLU(inplace=True), nn.Linear(feat_dim, feat_dim // 2))
        self._rounded = _RoundedWeights()

    # parameter order used by the autograd function
    def _params(self):
        ps = [self.x_embedding1.weight, self.x_embedding2.weight]
        for g, bn in zip(self.gnns, self.batch_norms):
            ps += [g.mlp[0].weight, g.mlp[0].bias, g.mlp[2].weight, g.mlp[2].bias,
                   g.edge_embedding1.weight, g.edge_embedding2.weight, bn.weight, bn.bias]
        ps += [self.feat_lin.weight, self.feat_lin.bias, self.out_lin[0].weight, self.out_lin[0].bias,
               self.out_lin[2].weight, self.out_lin[2].bias]
        return ps

    def forward(self, data):
        self._check_input(data)
        plan = get_plan(data)
        h, out = _GINetFunction.apply(self, plan, *self._params())
        return h, out

    def forward_pair(self, xis, xjs, grad_sink=None):
        """``(model(xis), model(xjs))`` -- the two encoder passes of ``MolCLR._step`` (molclr.py:57-60), in that order -- as ONE
        autograd node (see ``_GINetPairFunction``).  Same values as two separate calls."""
        self._check_input(xis)
        h_i, out_i, h_j, out_j = _GINetPairFunction.apply(self, get_plan(xis), get_plan(xjs), grad_sink, *self._params())
        return (h_i, out_i), (h_j, out_j)


def _lo(x, comp):
    return x if comp else None


def _head_forward(m, p, p_lo, rw, comp):
    """h = feat_lin(p); out = out_lin(h)   (ginet_molclr.py:114-115).  Returns h, out and what backward needs."""
    G, D, Fd = p.shape[0], m.emb_dim, m.feat_dim
    dev = p.device
    (Wf, Wf_lo), (W0, W0_lo), (W2, W2_lo) = rw.get(m.feat_lin.weight), rw.get(m.out_lin[0].weight), rw.get(m.out_lin[2].weight)
    h = torch.empty(G, Fd, device=dev)
    h_r = torch.empty(G, Fd, device=dev)
    h_lo = torch.empty(G, Fd, device=dev) if comp else None
    ops.gemm(p, Wf, G, Fd, D, A_lo=p_lo, B_lo=_lo(Wf_lo, comp), out=h, out2=h_r, out_lo=h_lo, bias=m.feat_lin.bias.detach())
    r = torch.empty(G, Fd, device=dev)
    r_lo = torch.empty(G, Fd, device=dev) if comp else None
    ops.gemm(h_r, W0, G, Fd, Fd, A_lo=h_lo, B_lo=_lo(W0_lo, comp), out=r, out_lo=r_lo, bias=m.out_lin[0].bias.detach(),
             relu=True, round_out=True)
    out = torch.empty(G, Fd // 2, device=dev)
    ops.gemm(r, W2, G, Fd // 2, Fd, A_lo=r_lo, B_lo=_lo(W2_lo, comp), out=out, bias=m.out_lin[2].bias.detach())
    return h, out, (h_r, r, Wf, W0, W2)


def _head_backward(m, p, saved, g_h, g_out):
    """Returns g_p and the six head parameter gradients (autograd of ginet_molclr.py:90-96,114-115)."""
    h_r, r, Wf, W0, W2 = saved
    G, D, Fd = p.shape[0], m.emb_dim, m.feat_dim
    dev = p.device
    g_out = g_out.contiguous()
    g_out_r = ops.round_tf32(g_out)
    dW2 = ops.gemm_dw(g_out_r, r, ordered=m.deterministic)
    db2 = ops.colsum(g_out)
    T = ops.colstat_tiles(G)
    # g_r = (g_out W2) * [r > 0]
    g_r = torch.empty(G, Fd, device=dev)
    part = torch.empty(T, Fd, device=dev)
    ops.gemm(g_out_r, W2, G, Fd, Fd // 2, b_mn=True, out=g_r, mask=r, round_out=True, colstat=part, colstat_mode=1)
    db0 = ops.reduce_partials(part, T, Fd, torch.empty(Fd, device=dev))
    dW0 = ops.gemm_dw(g_r, h_r, ordered=m.deterministic)
    # g_h(total) = g_r W0 (+ the gradient arriving on the returned representation h)
    g_hh_r = torch.empty(G, Fd, device=dev)
    part2 = torch.empty(T, Fd, device=dev)
    ops.gemm(g_r, W0, G, Fd, Fd, b_mn=True, out2=g_hh_r, addend=None if g_h is None else g_h.contiguous(),
             colstat=part2, colstat_mode=1)
    dbf = ops.reduce_partials(part2, T, Fd, torch.empty(Fd, device=dev))
    dWf = ops.gemm_dw(g_hh_r, p, ordered=m.deterministic)
    g_p = torch.empty(G, D, device=dev)
    ops.gemm(g_hh_r, Wf, G, D, Fd, b_mn=True, out=g_p)
    return g_p, (dWf, dbf, dW0, db0, dW2, db2)


def _encoder_forward(m, plan, comp, training, pool_mode):
    """Node embedding -> L x (aggregate, MLP, BatchNorm statistics) -> pooled graph vectors (ginet_molclr.py:103-113): ONE native
    call (csrc/gin_step.cu issues the kernel sequence listed in the module docstring).  Returns (p, p_lo, context): the pooled
    operand pair and what ``_encoder_backward`` needs.  Used by every GIN-E model of the package (pre-train and fine-tune heads)."""
    e = native.encoder_forward(m, plan, comp, training, pool_mode, with_head=hasattr(m, "out_lin"))
    return e.p, e.p_lo, e


def _encoder_backward(m, plan, e, g_p, training, pool_mode):
    """Backward of ``_encoder_forward`` given the gradient of the pooled vectors.  Returns the gradients of
    [x_embedding1, x_embedding2] + per layer [mlp0.w, mlp0.b, mlp2.w, mlp2.b, edge_emb1, edge_emb2, bn.w, bn.b] (views of one
    flat buffer)."""
    flat = native.grad_buffer(m, e)
    native.encoder_backward(m, e, g_p, flat, m.deterministic)
    views, _ = native.grad_views(m, e, flat, 0, 2 + 8 * m.num_layer)
    return views


def _check_precision(m):
    if m.precision not in PRECISIONS:
        raise ValueError(f"molclr_b200: precision must be one of {PRECISIONS}, got {m.precision!r}")
    return {"fp16x3": 2, "tf32x3": 1, "tf32": 0}[m.precision]      # = `comp` of the native calls


class _GINetFunction(torch.autograd.Function):
    """One view: encoder + projection head, forward and backward as two native calls each."""

    @staticmethod
    def forward(ctx, m, plan, *params):
        comp = _check_precision(m)
        m._refresh_weights(comp)
        e = native.encoder_forward(m, plan, comp, m.training, ops.POOL_MODES[m.pool_name])
        h, out = native.proj_head_forward(m, e)
        ctx.m, ctx.e = m, e
        return h, out

    @staticmethod
    def backward(ctx, g_h, g_out):
        m, e = ctx.m, ctx.e
        flat = native.grad_buffer(m, e)
        _view_backward(m, e, g_h, g_out, flat)
        ctx.e = None
        views, _ = native.grad_views(m, e, flat, 0, 2 + 8 * m.num_layer + 6)
        return (None, None, *views)


def _view_backward(m, e, g_h, g_out, flat, on_layer_done=None):
    if g_out is None:
        g_out = torch.zeros(e.plan.G, m.feat_dim // 2, device=e.dev)
    native.proj_head_backward(m, e, g_h, g_out, flat, m.deterministic)
    native.encoder_backward(m, e, None, flat, m.deterministic, on_layer_done)


class _GINetPairFunction(torch.autograd.Function):
    """Both views of a pre-training step (molclr.py:57-60) in one autograd node: the weight shadows are derived once, the two
    backward passes fill two flat gradient buffers that are summed slice by slice (no per-parameter accumulation kernels), and a
    data-parallel caller can start the all-reduce of a slice as soon as the second view's backward has produced it (``sink``)."""

    @staticmethod
    def forward(ctx, m, plan_i, plan_j, sink, *params):
        comp = _check_precision(m)
        m._refresh_weights(comp)
        pool_mode = ops.POOL_MODES[m.pool_name]
        e_i = native.encoder_forward(m, plan_i, comp, m.training, pool_mode)        # view i first: the order of the two
        h_i, out_i = native.proj_head_forward(m, e_i)                                # BatchNorm running-statistics updates
        e_j = native.encoder_forward(m, plan_j, comp, m.training, pool_mode)
        h_j, out_j = native.proj_head_forward(m, e_j)
        ctx.m, ctx.e_i, ctx.e_j, ctx.sink = m, e_i, e_j, sink
        return h_i, out_i, h_j, out_j

    @staticmethod
    def backward(ctx, g_hi, g_oi, g_hj, g_oj):
        m, e_i, e_j, sink = ctx.m, ctx.e_i, ctx.e_j, ctx.sink
        flat, flat2 = native.grad_buffer(m, e_j), native.grad_buffer(m, e_i)
        _view_backward(m, e_j, g_hj, g_oj, flat)
        sl = native.grad_slices(m, e_i)

        def merge(lo, hi, name):
            native.add_inplace(flat[lo:hi], flat2[lo:hi])
            if sink is not None:
                sink.slice_ready(flat, lo, hi, name)
        if g_oi is None:
            g_oi = torch.zeros(e_i.plan.G, m.feat_dim // 2, device=e_i.dev)
        native.proj_head_backward(m, e_i, g_hi, g_oi, flat2, m.deterministic)
        merge(*sl["head"], "head")
        native.encoder_backward(m, e_i, None, flat2, m.deterministic,
                                lambda l: merge(*(sl["embed"] if l < 0 else sl["layer"][l]), "embed" if l < 0 else f"layer{l}"))
        ctx.e_i = ctx.e_j = None
        views, _ = native.grad_views(m, e_i, flat, 0, 2 + 8 * m.num_layer + 6)
        if sink is not None:
            sink.backward_done(flat, views)
        return (None, None, None, None, *views)
