"""GINet (fine-tune variant): drop-in for ``models/ginet_finetune.py`` -- the same GIN-E encoder as the pre-training model
followed by ``feat_lin`` and the ``pred_head`` MLP (Linear -> Softplus/ReLU -> ... -> Linear(., 2 | 1)); ``forward(data)``
returns ``(h, pred_head(h))`` (ginet_finetune.py:129-147).  Encoder forward/backward are the kernels of ``ginet.py``;
the head runs on the same tcgen05 GEMM with small activation kernels.  The task loss (CrossEntropy / MSE / L1,
finetune.py:71-77) is the caller's.
"""
import torch
from torch import nn

from . import ops
from .ginet import (GINEConv, _EncoderBase, _RoundedWeights, _check_precision, _encoder_backward, _encoder_forward, _lo,
                    num_atom_type, num_chirality_tag)
from .graph import get_plan


class GINet(_EncoderBase):
    """ginet_finetune.py:52-147.

    Args:
        task (str): 'classification' (2 logits) | 'regression' (1 output)
        num_layer, emb_dim, feat_dim, drop_ratio, pool: as the pre-training model
        pred_n_layer (int): number of hidden layers of the prediction head (>= 1)
        pred_act (str): 'softplus' | 'relu'
    """

    def __init__(self, task="classification", num_layer=5, emb_dim=300, feat_dim=512, drop_ratio=0, pool="mean",
                 pred_n_layer=2, pred_act="softplus"):
        super().__init__()
        self.num_layer, self.emb_dim, self.feat_dim, self.drop_ratio, self.task = num_layer, emb_dim, feat_dim, drop_ratio, task
        if feat_dim % 8 != 0 or emb_dim % 4 != 0:
            # feat_dim // 2 is the width of the head's hidden activations -- tensor-core operands, whose widths are multiples of 4
            raise ValueError(f"molclr_b200: feat_dim must be a multiple of 8 and emb_dim a multiple of 4, got feat_dim={feat_dim}, emb_dim={emb_dim}")
        self.pool_name = pool
        self.x_embedding1 = nn.Embedding(num_atom_type, emb_dim)
        self.x_embedding2 = nn.Embedding(num_chirality_tag, emb_dim)
        nn.init.xavier_uniform_(self.x_embedding1.weight.data)
        nn.init.xavier_uniform_(self.x_embedding2.weight.data)
        self.gnns = nn.ModuleList([GINEConv(emb_dim) for _ in range(num_layer)])
        self.batch_norms = nn.ModuleList([nn.BatchNorm1d(emb_dim) for _ in range(num_layer)])
        self.feat_lin = nn.Linear(emb_dim, feat_dim)
        if task == "classification":
            out_dim = 2
        elif task == "regression":
            out_dim = 1
        else:                              # the reference hits an unbound `out_dim` here (ginet_finetune.py:96-99)
            raise UnboundLocalError("local variable 'out_dim' referenced before assignment")
        self.pred_n_layer = max(1, pred_n_layer)
        if pred_act == "relu":
            act = lambda: nn.ReLU(inplace=True)
        elif pred_act == "softplus":
            act = nn.Softplus
        else:
            raise ValueError("Undefined activation function")                         # ginet_finetune.py:123-124
        self.pred_act = pred_act
        head = [nn.Linear(feat_dim, feat_dim // 2), act()]
        for _ in range(self.pred_n_layer - 1):
            head.extend([nn.Linear(feat_dim // 2, feat_dim // 2), act()])
        head.append(nn.Linear(feat_dim // 2, out_dim))
        self.pred_head = nn.Sequential(*head)
        self._rounded = _RoundedWeights()

    def _head_linears(self):
        return [mod for mod in self.pred_head if isinstance(mod, nn.Linear)]

    def _params(self):
        ps = [self.x_embedding1.weight, self.x_embedding2.weight]
        for g, bn in zip(self.gnns, self.batch_norms):
            ps += [g.mlp[0].weight, g.mlp[0].bias, g.mlp[2].weight, g.mlp[2].bias,
                   g.edge_embedding1.weight, g.edge_embedding2.weight, bn.weight, bn.bias]
        ps += [self.feat_lin.weight, self.feat_lin.bias]
        for lin in self._head_linears():
            ps += [lin.weight, lin.bias]
        return ps

    def forward(self, data):
        self._check_input(data)
        plan = get_plan(data)
        return _FinetuneFunction.apply(self, plan, *self._params())

    def load_my_state_dict(self, state_dict):
        """ginet_finetune.py:149-157: copy the entries whose names exist here (pre-trained encoder -> fine-tune model)."""
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name not in own_state:
                continue
            if isinstance(param, nn.parameter.Parameter):
                param = param.data
            own_state[name].copy_(param)


def _pad4(w_hi, w_lo, rows, dev):
    """Zero-pads a [rows < 4, K] weight pair to 4 rows (GEMM output widths are multiples of 4)."""
    if rows % 4 == 0:
        return w_hi, w_lo
    r4 = (rows + 3) // 4 * 4
    out = []
    for w in (w_hi, w_lo):
        buf = torch.zeros(r4, w.stride(0), device=dev)
        ops.copy_rows(w, buf, rows)
        out.append(buf[:, :w.shape[1]])
    return out


def finetune_head_forward(m, p, p_lo, lins, mode, comp):
    """feat_lin followed by the prediction MLP ``lins`` (activation ``mode`` between the Linears): ginet_finetune.py:144-147,
    gcn_finetune.py:158-160.  Returns (h, pred, saved, Wf)."""
    G, D, Fd = p.shape[0], m.emb_dim, m.feat_dim
    dev, rw = p.device, m._rounded
    Wf, Wf_lo = rw.get(m.feat_lin.weight)
    h, h_r = torch.empty(G, Fd, device=dev), torch.empty(G, Fd, device=dev)
    h_lo = torch.empty(G, Fd, device=dev) if comp else None
    ops.gemm(p, Wf, G, Fd, D, A_lo=p_lo, B_lo=_lo(Wf_lo, comp), out=h, out2=h_r, out_lo=h_lo, bias=m.feat_lin.bias.detach())
    x_hi, x_lo, saved = h_r, h_lo, []
    for k, lin in enumerate(lins):
        O, I = lin.weight.shape
        W, W_lo = rw.get(lin.weight)
        last = k == len(lins) - 1
        O4 = (O + 3) // 4 * 4
        bias = lin.bias.detach()
        if O4 != O:
            W, W_lo = _pad4(W, W_lo, O, dev)
            bias4 = torch.zeros(O4, device=dev)
            ops.copy_rows(bias.reshape(O, 1), bias4.reshape(O4, 1), O)
            bias = bias4
        t = torch.empty(G, O4, device=dev)
        ops.gemm(x_hi, W, G, O4, I, A_lo=x_lo, B_lo=_lo(W_lo, comp), out=t, bias=bias)
        saved.append((x_hi, t, W, O))
        if not last:
            x_hi, x_lo = ops.act_fwd(t, mode, comp)
    return h, t[:, :lins[-1].weight.shape[0]], saved, Wf


def finetune_head_backward(m, p, saved, Wf, g_h, g_pred, mode):
    """Returns (g_p, dWf, dbf, head_grads) with head_grads = [dW, db] per Linear of the prediction MLP."""
    G, D, Fd = p.shape[0], m.emb_dim, m.feat_dim
    dev = p.device
    head_grads = [None] * (2 * len(saved))
    # gradient arriving on the (4-padded) head output
    x_hi, t, W, O = saved[-1]
    g_t = torch.zeros(G, t.shape[1], device=dev)
    if g_pred is not None:
        ops.copy_cols(g_pred.contiguous(), g_t, O)
    g_t = ops.round_tf32(g_t)
    g_x = None
    for k in range(len(saved) - 1, -1, -1):
        x_hi, t, W, O = saved[k]
        I = x_hi.shape[1]
        dW = ops.gemm_dw(g_t, x_hi, ordered=m.deterministic)                          # [O4, I]
        head_grads[2 * k], head_grads[2 * k + 1] = dW[:O], ops.colsum(g_t)[:O]
        g_x = torch.empty(G, I, device=dev)
        if k > 0:
            ops.gemm(g_t, W, G, I, t.shape[1], b_mn=True, out=g_x)
            g_t = ops.act_bwd(g_x, saved[k - 1][1], mode)    # through the activation in front of this Linear
        else:                                                # reaches h: add the gradient arriving on the returned h
            g_hh_r = torch.empty(G, I, device=dev)
            ops.gemm(g_t, W, G, I, t.shape[1], b_mn=True, out=g_x, out2=g_hh_r, addend=None if g_h is None else g_h.contiguous())
    dWf = ops.gemm_dw(g_hh_r, p, ordered=m.deterministic)
    dbf = ops.colsum(g_x)
    g_p = torch.empty(G, D, device=dev)
    ops.gemm(g_hh_r, Wf, G, D, Fd, b_mn=True, out=g_p)
    return g_p, dWf, dbf, head_grads


class _FinetuneFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, m, plan, *params):
        comp = _check_precision(m)
        training, pool_mode = m.training, ops.POOL_MODES[m.pool_name]
        m._refresh_weights(comp)
        p, p_lo, layers = _encoder_forward(m, plan, comp, training, pool_mode)
        mode = ops.ACT_MODES[m.pred_act]
        h, pred, saved, Wf = finetune_head_forward(m, p, p_lo, m._head_linears(), mode, comp)
        ctx.m, ctx.plan, ctx.layers, ctx.p, ctx.saved, ctx.Wf = m, plan, layers, p, saved, Wf
        ctx.training, ctx.pool_mode, ctx.mode = training, pool_mode, mode
        return h, pred

    @staticmethod
    def backward(ctx, g_h, g_pred):
        m, plan, p = ctx.m, ctx.plan, ctx.p
        g_p, dWf, dbf, head_grads = finetune_head_backward(m, p, ctx.saved, ctx.Wf, g_h, g_pred, ctx.mode)
        grads = _encoder_backward(m, plan, ctx.layers, g_p, ctx.training, ctx.pool_mode)
        ctx.layers = None
        return (None, None, *grads, dWf, dbf, *head_grads)
