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

and the backward mirrors it (see ``_backward``).  Nothing here falls back to PyTorch operators.
"""
import torch
from torch import nn

from . import ops
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
    by ONE kernel launch at the start of every forward (``refresh``).  Nothing is cached across forwards: in-place parameter
    updates (``torch.optim.Adam(fused=True)``, ``param.data`` writes, ``dist.broadcast``) do not bump tensor version counters,
    so any cache keyed on them serves stale weights."""

    def __init__(self):
        self._cur = {}

    def refresh(self, specs):
        """specs: [(parameter, ops.W_* flags)].  The shadows of exactly these parameters are valid until the next refresh."""
        outs = ops.prepare_weights([(p.detach(), f) for p, f in specs])
        self._cur = {id(p): (p.data_ptr(), o) for (p, _), o in zip(specs, outs)}

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

    def b16(self, p):
        """bf16 correction tiles [2, rows16, ld16] of ``raw(p)`` (None unless part of the last refresh)."""
        e = self._entry(p, "b16")
        return None if e is None else e["b16"]


PRECISIONS = ("tf32x3", "tf32")


class _EncoderBase(nn.Module):
    """Shared plumbing of the GINet / GCN drop-ins.

    ``precision`` (attribute, not a constructor argument -- the constructor is the reference's):
      * ``"tf32x3"`` (default): every FORWARD contraction is the error-compensated 3-pass TF32 product
        (~fp32 accuracy), so pre-activations -- and with them the ReLU masks the backward pass depends on --
        match the fp32 reference; BACKWARD contractions are single-pass TF32.
      * ``"tf32"``: single-pass TF32 everywhere (fastest; activations carry ~1e-3 relative error and the
        resulting ReLU mask flips show up as percent-level noise in gradients).
    """
    precision = "tf32x3"

    def _check_input(self, data):
        if not 0 <= self.drop_ratio < 1:
            raise ValueError(f"dropout probability has to be between 0 and 1, but got {self.drop_ratio}")
        if self.pool_name not in ops.POOL_MODES:
            # the reference leaves self.pool unset for unknown names and fails at forward (ginet_molclr.py:83-88,113)
            raise AttributeError(f"'{type(self).__name__}' object has no attribute 'pool'")

    deterministic = False      # True: weight gradients are summed in a fixed order (bit-reproducible run to run); see DESIGN.md

    def _gemm_weights(self, comp):
        """(parameter, ops.W_* operand forms) of every contraction of the model.  GINEConv MLP weights / GCNConv weights: hi
        (backward dX products, single-pass forward) and, for the compensated forward, the unrounded K-major copy + its bf16
        correction tiles; every other nn.Linear (projection / prediction heads): hi (+ lo for the explicit 3-pass product)."""
        enc = ops.W_HI | (ops.W_B16 if comp else 0)
        head = ops.W_HI | (ops.W_LO if comp else 0)
        specs, seen = [], set()
        for g in self.gnns:
            if hasattr(g, "mlp"):
                ws = [(g.mlp[0].weight, enc | (ops.W_RAW if comp else 0)), (g.mlp[2].weight, enc | (ops.W_RAW if comp else 0))]
            else:                                         # GCNConv: stored [in, out]
                ws = [(g.weight, enc | (ops.W_RAW_T if comp else 0))]
            specs += ws
            seen.update(id(w) for w, _ in ws)
        for mod in self.modules():
            if isinstance(mod, nn.Linear) and id(mod.weight) not in seen:
                seen.add(id(mod.weight))
                specs.append((mod.weight, head))
        return specs

    def _refresh_weights(self, comp):
        self._rounded.refresh(self._gemm_weights(comp))

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
    """Node embedding -> L x (aggregate, MLP, BatchNorm statistics) -> pooled graph vectors (ginet_molclr.py:103-113).
    Returns (p, p_lo, layers): the pooled operand pair and the per-layer tensors the backward needs."""
    L, D, H, N = m.num_layer, m.emb_dim, 2 * m.emb_dim, plan.N
    dev = m.x_embedding1.weight.device
    rw = m._rounded
    h0 = ops.embed_nodes_fwd(plan, m.x_embedding1.weight.detach(), m.x_embedding2.weight.detach())
    src, coef_prev = h0, None
    layers = []
    T = ops.colstat_tiles(N)
    drops = m._dropout_seeds()                # drops[l]: dropout applied to layer l's output (ginet_molclr.py:108-111)
    for l in range(L):
        g, bn = m.gnns[l], m.batch_norms[l]
        dp = drops[l - 1] if l > 0 else (0, 0.0)
        # a and u stay UNROUNDED fp32 (one tensor each): the compensated GEMMs derive their low halves on chip
        a = ops.gine_aggregate_fwd(plan, src, g.edge_embedding1.weight.detach(), g.edge_embedding2.weight.detach(),
                                   bn_coef=coef_prev, relu=True, round_out=False, drop=dp)
        # W1 / W2: tf32-rounded copies (single-pass forward, and the backward's dX GEMMs); tf32x3: the raw weights, the
        # compensated GEMM derives the bf16 correction tiles of both operands on chip
        (W1, _), (W2, _) = rw.get(g.mlp[0].weight), rw.get(g.mlp[2].weight)
        B1, B2 = (rw.raw(g.mlp[0].weight), rw.raw(g.mlp[2].weight)) if comp else (W1, W2)
        S1, S2 = (rw.b16(g.mlp[0].weight), rw.b16(g.mlp[2].weight)) if comp else (None, None)
        u = ops.padded(N, H, dev)
        ubits = ops.relu_bits_buffer(N, H, dev)       # [u > 0] as bits: the backward GEMM's mask (8 MB instead of 246)
        ops.gemm(a, B1, N, H, D, compensate=comp, B16=S1, out=u, bias=g.mlp[0].bias.detach(), relu=True, relu_bits=ubits)
        z = torch.empty(N, D, device=dev)
        if training:
            stats = torch.empty(T, 2, D, device=dev)
            ops.gemm(u, B2, N, D, H, compensate=comp, B16=S2, out=z, bias=g.mlp[2].bias.detach(), colstat=stats, colstat_mode=2)
            momentum = 0.1 if bn.momentum is None else bn.momentum
            coef = ops.bn_fwd_finalize(stats, T, N, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                       bn.num_batches_tracked, momentum, bn.eps)
        else:
            ops.gemm(u, B2, N, D, H, compensate=comp, B16=S2, out=z, bias=g.mlp[2].bias.detach())
            coef = ops.bn_eval_coef(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
        layers.append((a, u, z, coef, W1, W2, ubits))
        src, coef_prev = z, coef
    argmax = torch.empty(plan.G, D, dtype=torch.int32, device=dev) if pool_mode == 2 else None
    p, p_lo = ops.pool_fwd(plan, src, coef_prev, pool_mode, relu=False, round_out=True, want_lo=True, argmax=argmax, drop=drops[L - 1]) \
        if comp else (ops.pool_fwd(plan, src, coef_prev, pool_mode, relu=False, round_out=True, argmax=argmax, drop=drops[L - 1]), None)
    layers.append((drops, argmax))            # trailing entry: what the backward needs besides the per-layer tensors
    return p, p_lo, layers


def _encoder_backward(m, plan, layers, g_p, training, pool_mode):
    """Backward of ``_encoder_forward`` given the gradient of the pooled vectors.  Returns the gradients of
    [x_embedding1, x_embedding2] + per layer [mlp0.w, mlp0.b, mlp2.w, mlp2.b, edge_emb1, edge_emb2, bn.w, bn.b]."""
    L, D, H, N = m.num_layer, m.emb_dim, 2 * m.emb_dim, plan.N
    dev = g_p.device
    grads = [None] * (2 + 8 * L)
    drops, argmax = layers[L]
    # last layer: BatchNorm backward fed by the pool backward (g_y is never materialised)
    a, u, z, coef, W1, W2, ubits = layers[L - 1]
    bn = m.batch_norms[L - 1]
    partials, P = ops.pool_bwd_stats(plan, g_p, z, coef, pool_mode, argmax=argmax, drop=drops[L - 1])
    dgamma, dbeta, bcoef = ops.bn_bwd_finalize(partials, P, N, bn.weight.detach(), coef, training)
    g_z, db2 = ops.bn_bwd_apply(z, bcoef, gp=g_p, plan=plan, pool_mode=pool_mode, argmax=argmax, drop=drops[L - 1])
    T = ops.colstat_tiles(N)
    for l in range(L - 1, -1, -1):
        a, u, z, coef, W1, W2, ubits = layers[l]
        base = 2 + 8 * l
        grads[base + 6], grads[base + 7], grads[base + 3] = dgamma, dbeta, db2
        # g_u = (g_z W2) * [u > 0];  db1 = colsum(g_u)
        g_u = ops.padded(N, H, dev)
        part = torch.empty(T, H, device=dev)
        ops.gemm(g_z, W2, N, H, D, b_mn=True, out=g_u, mask_bits=ubits, round_out=True, colstat=part, colstat_mode=1)
        grads[base + 1] = ops.reduce_partials(part, T, H, torch.empty(H, device=dev))
        grads[base + 2] = ops.gemm_dw(g_z, u, ordered=m.deterministic)                 # dW2 [D, H]
        g_a = torch.empty(N, D, device=dev)
        ops.gemm(g_u, W1, N, D, H, b_mn=True, out=g_a)
        grads[base + 0] = ops.gemm_dw(g_u, a, ordered=m.deterministic)                 # dW1 [H, D]
        grads[base + 4], grads[base + 5] = ops.edge_table_grad(plan, g_a)
        if l > 0:
            _, _, zp, coefp, _, _, _ = layers[l - 1]
            bnp = m.batch_norms[l - 1]
            g_y, partials, P = ops.gine_aggregate_bwd(plan, g_a, z_prev=zp, bn_coef=coefp, relu=True, drop=drops[l - 1])
            dgamma, dbeta, bcoef = ops.bn_bwd_finalize(partials, P, N, bnp.weight.detach(), coefp, training)
            g_z, db2 = ops.bn_bwd_apply(zp, bcoef, gy=g_y)
        else:
            g_h0, _, _ = ops.gine_aggregate_bwd(plan, g_a)
            grads[0], grads[1] = ops.embed_nodes_bwd(plan, g_h0)
    return grads


def _check_precision(m):
    if m.precision not in PRECISIONS:
        raise ValueError(f"molclr_b200: precision must be one of {PRECISIONS}, got {m.precision!r}")
    return m.precision == "tf32x3"


class _GINetFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, m, plan, *params):
        comp = _check_precision(m)
        training = m.training
        pool_mode = ops.POOL_MODES[m.pool_name]
        m._refresh_weights(comp)
        p, p_lo, layers = _encoder_forward(m, plan, comp, training, pool_mode)
        h, out, head_saved = _head_forward(m, p, p_lo, m._rounded, comp)
        ctx.m, ctx.plan, ctx.layers, ctx.p, ctx.head_saved = m, plan, layers, p, head_saved
        ctx.training, ctx.pool_mode = training, pool_mode
        return h, out

    @staticmethod
    def backward(ctx, g_h, g_out):
        m, plan, layers, p = ctx.m, ctx.plan, ctx.layers, ctx.p
        if g_out is None:
            g_out = torch.zeros(p.shape[0], m.feat_dim // 2, device=p.device)
        g_p, head_grads = _head_backward(m, p, ctx.head_saved, g_h, g_out)
        grads = _encoder_backward(m, plan, layers, g_p, ctx.training, ctx.pool_mode)
        ctx.layers = None
        return (None, None, *grads, *head_grads)
